#!/bin/bash
# A/B of the product library against a side build (tools/lib/$1) in the same box / same process order, alternating
cd "${GRAFT_REPO_ROOT:-/root/repo}"
SIDE=tools/lib/$1; shift
for rep in 1 2; do
  echo "--- product"; VARIANTS=8 timeout -s KILL 300 python tools/attn_ragged_probe.py ${PAIRS:-1026x1026,1024x1024}
  echo "--- side ($SIDE)"; PCD_B200_LIB=$PWD/$SIDE VARIANTS=8 timeout -s KILL 300 python tools/attn_ragged_probe.py ${PAIRS:-1026x1026,1024x1024}
done

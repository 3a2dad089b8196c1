#!/bin/bash
# A/B two builds of the library on the SAME box, alternating runs: tools/ab.sh <python args...>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for rep in 1 2; do
  for v in base new; do
    echo "--- $v (rep $rep)"
    PCD_B200_LIB=$PWD/ab/libpcd_$v.so python "$@" 2>&1 | tail -${ABTAIL:-6}
  done
done

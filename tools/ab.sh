#!/bin/bash
# A/B two builds of the library on the SAME box, alternating runs: tools/ab.sh <python args...>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for rep in $(seq 1 ${ABREPS:-2}); do
  for v in base new; do
    echo "--- $v (rep $rep)"
    PCD_B200_LIB=$PWD/ab/libpcd_$v.so timeout -s KILL ${ABTIMEOUT:-300} python "$@" 2>&1 | tail -${ABTAIL:-6}
  done
done

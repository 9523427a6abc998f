#!/bin/bash
# A/B build only (python .../build_ext.py --force with AGCN_AB_SWITCHES=1): the pipelined e2e loop for several pack grids
# (CTAs x threads); PACKTIME=1 also times the pack kernels alone
SHAPES=${SHAPES:-1x256 1x512 1x1024 2x512 2x1024 4x1024 8x1024}
for W in ${WORKLOADS:-C2}; do
for shape in $SHAPES; do
  c=${shape%x*}; t=${shape#*x}
  a=""
  [ -n "$PACKTIME" ] && a=$(AGCN_PACK_CTAS=$c AGCN_PACK_THREADS=$t timeout 200 python tools/pack_time.py $W 2>&1 | tail -2 | cut -c1-26 | paste -s -d' ')
  b=$(AGCN_PACK_CTAS=$c AGCN_PACK_THREADS=$t timeout 200 python tools/e2e_host_split.py $W 2>&1 | tail -2 | cut -c1-24 | head -1)
  echo "$W pack grid $shape: $a | e2e loop $b"
done
done

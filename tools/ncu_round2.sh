#!/bin/bash
# ncu evidence of round 2 (run on the GPU box from the repo root): launch list of the eager C2 step and full captures of
# the recurrence tiles / the all-rows contraction (C2), of the row-tiled tensor-core product (C3) and of the metric block of
# big graphs (C3, paper semantics, full metric gradient).  Outputs under gpurun_out/.  PARTS="c2 c3 c3paper" selects.
mkdir -p gpurun_out
PARTS=${PARTS:-"c2 c3 c3paper"}
B="python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --no-paper --no-configs"
[[ " $PARTS " == *" c2 "* ]] && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_c2.csv $B > gpurun_out/r02_ncu_launches.log 2>&1
[[ " $PARTS " == *" c2 "* ]] && timeout 500 ncu --set full --clock-control none --import-source on -k regex:"cheb_tile_fwd_kernel|rows_gemm_kernel|tc_gemm_tn_kernel" -s 30 -c 12 -f -o gpurun_out/r02_ncu_full_c2 $B > gpurun_out/r02_ncu_full_c2.log 2>&1
[[ " $PARTS " == *" c3 "* ]] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"grouped_tc|rows_gemm_kernel" -s 30 -c 6 -f -o gpurun_out/r02_ncu_full_c3 python bench.py --workload C3 --steps 2 --warmup 3 --skip-cpu --no-graph --no-paper --no-configs > gpurun_out/r02_ncu_full_c3.log 2>&1
[[ " $PARTS " == *" c3paper "* ]] && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pair_tcu_kernel|big_trans_kernel|big_sweep_kernel" -s 40 -c 12 -f -o gpurun_out/r02_ncu_full_c3_paper python tools/step_timeline.py --workload C3 --paper > gpurun_out/r02_ncu_full_c3_paper.log 2>&1
ls -la gpurun_out | grep r02_

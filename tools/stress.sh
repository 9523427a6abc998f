#!/bin/bash
# Thousands of back-to-back steps of every workload, eager and under CUDA-graph replay (rare races / protocol bugs of the
# hand-written kernels show up as a launch failure or a trap after ten seconds).  Outputs under gpurun_out/stress_*.log
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; timeout 300 "$@" > $O/stress_$name.log 2>&1 ; echo "$name rc $? $(tail -n 1 $O/stress_$name.log | cut -c1-260)"; }
run c2_replay python tools/stress_step.py --workload C2 --mode replay --iters 4000 --smi
run c2_eager python tools/stress_step.py --workload C2 --mode eager --iters 2500
run c1_replay python tools/stress_step.py --workload C1 --mode replay --iters 4000
run c3_replay python tools/stress_step.py --workload C3 --mode replay --iters 1500
run c4_replay python tools/stress_step.py --workload C4 --mode replay --iters 1500
run c2_paper_replay python tools/stress_step.py --workload C2 --mode replay --iters 1000 --paper

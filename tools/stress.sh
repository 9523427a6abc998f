#!/bin/bash
# Reproduce / localise rare failures of the resident step.  Outputs under gpurun_out/stress_*.log
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "== $name: $*" ; timeout 240 "$@" > $O/stress_$name.log 2>&1 ; echo "rc $?" ; tail -n 2 $O/stress_$name.log | cut -c1-600; }
run replay python tools/stress_step.py --mode replay --iters 4000
run eager python tools/stress_step.py --mode eager --iters 2500
run eager_noflush python tools/stress_step.py --mode eager --iters 2500 --no-flush
run blocking env CUDA_LAUNCH_BLOCKING=1 python tools/stress_step.py --mode eager --iters 1500 --no-flush

#!/bin/bash
# One GPU session: parity of the build (and of the SIMT fallback of the row-tiled products), timings of the
# C1/C3/C4/C5 layer shapes, bench line, ncu capture of grouped_tc_kernel.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_check.sh [quick]
mkdir -p gpurun_out
O=gpurun_out
run() { echo "== $*" ; timeout 170 "$@" ; echo "== rc $?" ; }
{
run python tools/dbg_grouped.py
run python -m pytest tests -m gpu -q 2>&1 | tail -25
run env AGCN_BIG_TC=0 python -m pytest tests -m gpu -q 2>&1 | tail -8
} > $O/parity.log 2>&1
grep -E "passed|failed|worst|rc " $O/parity.log
rm -f $O/layer_sweep.jsonl
timeout 150 python tools/layer_sweep.py --group c3,c4,c1 --iters 5 --budget-s 90 --out $O/layer_sweep.jsonl --tag r01_m > /dev/null 2>> $O/sweep.err
timeout 200 python tools/layer_sweep.py --group sweep,sweep_big --match "K=3" --iters 5 --budget-s 120 --out $O/layer_sweep.jsonl --tag r01_m > /dev/null 2>> $O/sweep.err
wc -l $O/layer_sweep.jsonl
if [ "$1" != "quick" ]; then
timeout 240 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --no-graph --no-paper > $O/ncu_bench.log 2>&1
fi
timeout 150 ncu --set full --clock-control none --import-source on -k regex:grouped_tc -s 6 -c 2 -o $O/ncu_grouped_tc -f \
  python tools/layer_sweep.py --group c3 --match "F=128 ,literal" --iters 2 --out $O/ncu_sweep.jsonl > $O/ncu.log 2>&1
ls -la $O | tail -20

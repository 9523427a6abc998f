#!/bin/bash
# One GPU session: parity of the default build and of the opt-in big-graph paths, A/B timings, bench, ncu.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_check.sh
mkdir -p gpurun_out
O=gpurun_out
run() { echo "== $*" ; timeout 170 "$@" ; echo "== rc $?" ; }
{
run python tools/dbg_grouped.py
run env AGCN_BIG_TC=1 python -m pytest tests -m gpu -q 2>&1 | tail -40
run python -m pytest tests -m gpu -q 2>&1 | tail -15
run env AGCN_BIG_TC=1 AGCN_CHEB_SMALL_MAX=64 python -m pytest tests -m gpu -q 2>&1 | tail -25
} > $O/parity.log 2>&1
tail -5 $O/parity.log
rm -f $O/ab_sweep.jsonl
AGCN_BIG_TC=0 timeout 120 python tools/layer_sweep.py --group c3,c4 --iters 5 --budget-s 60 --out $O/ab_sweep.jsonl --tag simt > /dev/null 2>> $O/sweep.err
AGCN_BIG_TC=1 timeout 120 python tools/layer_sweep.py --group c3,c4 --iters 5 --budget-s 60 --out $O/ab_sweep.jsonl --tag big_tc > /dev/null 2>> $O/sweep.err
AGCN_BIG_TC=1 timeout 120 python tools/layer_sweep.py --group sweep,sweep_big --match "K=3" --iters 5 --budget-s 80 --out $O/ab_sweep.jsonl --tag big_tc > /dev/null 2>> $O/sweep.err
AGCN_BIG_TC=1 AGCN_CHEB_SMALL_MAX=64 timeout 100 python tools/layer_sweep.py --group c1,sweep --match "N=128 " --iters 5 --budget-s 50 --out $O/ab_sweep.jsonl --tag big_tc_mid64 > /dev/null 2>> $O/sweep.err
AGCN_BIG_TC=1 AGCN_CHEB_SMALL_MAX=64 timeout 100 python tools/layer_sweep.py --group c1 --iters 5 --budget-s 40 --out $O/ab_sweep.jsonl --tag big_tc_mid64 > /dev/null 2>> $O/sweep.err
wc -l $O/ab_sweep.jsonl
timeout 200 python bench.py > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json
AGCN_BIG_TC=1 AGCN_CHEB_SMALL_MAX=64 timeout 100 python bench.py --skip-cpu --no-paper > $O/bench_mid64.json 2> $O/bench_mid64.err
AGCN_BIG_TC=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:grouped_tc -s 6 -c 2 -o $O/ncu_grouped_tc -f \
  python tools/layer_sweep.py --group c3 --match "F=128 ,literal" --iters 2 --out $O/ncu_sweep.jsonl > $O/ncu.log 2>&1
ls -la $O | tail -20

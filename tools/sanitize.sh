#!/bin/bash
# compute-sanitizer over a subset of the GPU tests: memcheck, synccheck, racecheck (shared-memory
# hazards of the hand-rolled mbarrier / TMEM protocols).  Summaries land in gpurun_out/sanitizer_*.log.
# Usage (on the GPU box, from the repo root):  bash tools/sanitize.sh
mkdir -p gpurun_out
rm -f gpurun_out/sanitizer_summary.log
SEL='test_sgc_ll_forward_backward or test_reslap_forward_backward or test_feature_shapes or test_big_graphs_all_modes or test_head_loss or test_first_layer_backward or test_equal_size_big_graphs_metric_block or test_uniform_tensor_core_product'
SEL_SMALL='test_sgc_ll_forward_backward or test_feature_shapes or test_equal_size_big_graphs_metric_block'
for tool in memcheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_block_layers.py -m gpu -q -x -k "$SEL or block or mlp or dropout" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed|Error|RACECHECK SUMMARY" gpurun_out/sanitizer_$tool.log | tail -5 >> gpurun_out/sanitizer_summary.log
done
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_SMALL" > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_summary.log
grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_racecheck.log | tail -8 >> gpurun_out/sanitizer_summary.log
cat gpurun_out/sanitizer_summary.log

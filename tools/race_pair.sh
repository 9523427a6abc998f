#!/bin/bash
# racecheck of the tensor-core pair kernel alone (hazard report with thread ids)
timeout 600 compute-sanitizer --tool racecheck --racecheck-report ${REPORT:-analysis} --error-exitcode 7 --print-limit 6 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_equal_size_big_graphs_metric_block" > gpurun_out/sanitizer_racecheck_pair.log 2>&1; echo rc=$?; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/sanitizer_racecheck_pair.log | head

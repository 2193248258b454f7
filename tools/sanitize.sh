#!/bin/bash
# compute-sanitizer over the operator-level GPU tests and one tiny end-to-end engine test (run under gpurun).
#   bash tools/sanitize.sh [tag]  ->  gpurun_out/sanitize_<tool>_<tag>.log + a summary line per tool
# memcheck: out-of-bounds / misaligned accesses of every kernel incl. TMA-fed shared memory; racecheck: shared-memory hazards
# (mbarrier-protected producer/consumer hand-offs are the interesting part); synccheck: barrier misuse.
# The sanitizer slows kernels 10-100x: the selection keeps one case of every epilogue / kernel variant, small extents.
TAG=${1:-run}
mkdir -p gpurun_out
SEL_IGEMM='test_linear and (128-64-64 or 1000-640) or test_geglu or test_conv3x3 and (16-16-64 or 8-8-1280 or 24-24-128 or 320-4-1) or test_conv_concat or test_swapped_operands or test_conv3x3_swapped and (16-16-64 or 8-8-1280-1280-1-64) or test_tconv and (16-8 or 40-28) or test_linear_row_statistics and 256-1280 or test_linear_layernorm_folded and 64-1280 or test_geglu_layernorm'
SEL_OPS='not 4096'
for tool in memcheck racecheck synccheck; do
  log=gpurun_out/sanitize_${tool}_${TAG}.log
  : > $log
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_igemm_gpu.py -q -m gpu -x -k "$SEL_IGEMM" >> $log 2>&1
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "$SEL_OPS" >> $log 2>&1
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_engine_gpu.py -q -m gpu -x -k "test_tiny_stream_loop and True" >> $log 2>&1
  echo "== $tool: $(grep -E 'passed|failed' $log | tr '\n' ' ') | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tr '\n' ' ')"
done

#!/bin/bash
# GPU parity suite, one pytest process per file (a trapped kernel poisons only its own file); logs under gpurun_out/.
# usage: bash tools/gpu_tests.sh <tag> [file ...]
TAG=${1:-run}; shift
mkdir -p gpurun_out
FILES=${@:-$(ls tests/test_*gpu*.py)}
rc=0
for f in $FILES; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -x -s > gpurun_out/${n}_${TAG}.log 2>&1 || rc=1
  echo "$n: $(grep -E 'passed|failed|error' gpurun_out/${n}_${TAG}.log | tail -1)"
done
exit $rc

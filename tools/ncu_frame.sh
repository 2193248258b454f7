#!/bin/bash
# ncu evidence for one bench run (run under gpurun): launch list with durations + DRAM bytes for every kernel of ~1 frame.
# usage: bash tools/ncu_frame.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -s 1100 -c 520 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_bench_${TAG}.log | cut -c1-200
wc -l gpurun_out/launches_${TAG}.csv

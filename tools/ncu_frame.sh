#!/bin/bash
# ncu evidence for one bench run (run under gpurun): launch list with durations + DRAM bytes + tensor-pipe activity for every
# kernel of ~2 frames.  usage: bash tools/ncu_frame.sh <tag>   ->  gpurun_out/launches_<tag>.csv
# (blocking-call pattern, one lane running the default pipeline's throughput launch policy -- CTA pairs, 100 KB rings --: the
# launch list of ONE frame; numbers printed by a run under ncu are never bench values)
TAG=${1:-r02}
mkdir -p gpurun_out
B200SD_LANES=1 B200SD_POLICY_FRAMES=${POLICY_FRAMES:-8} ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'igemm_kernel|igemm_pair_kernel|tconv_kernel|attn_kernel|gn_|layernorm|smallconv_kernel|upsample2x|post_u8|lcm_step' \
    -c 1400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-library-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
tail -1 gpurun_out/ncu_bench_${TAG}.log | cut -c1-200
wc -l gpurun_out/launches_${TAG}.csv

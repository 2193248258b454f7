#!/bin/bash
# Round-2 evidence batch on one B200 (run under gpurun): tests, headline bench + baselines, other workloads, ncu captures.
TAG=${1:-r02h}
mkdir -p gpurun_out
bash tools/gpu_tests.sh $TAG
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_{tag}.json").readlines()[-1]); r = d["roofline"]
print(f"HEADLINE {d['value']:.1f} fps  {d['ms_per_step']:.3f} ms  seq {d['sequential']['ms_per_frame']:.3f} ms  e2e {d['e2e']['value']:.1f} p50 {d['e2e']['p50_ms']:.2f} p99 {d['e2e']['p99_ms']:.2f}  "
      f"igemm frac {r['frac']:.3f} ({r['kernel_ms_per_step']:.2f} ms)  lib {d['library_baseline']}  cpu {d['cpu_baseline']['value']:.3f} fps x{d['cpu_baseline']['cores']}  clocks {d['clocks']['sm_mhz']}/{d['clocks']['sm_max_mhz']} {d['clocks']['reasons']}")
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; cut -c1-220 gpurun_out/bench_ref_${TAG}.json
python bench.py --impl library --steps 100 --warmup 5 > gpurun_out/bench_lib_${TAG}.json 2>/dev/null; cut -c1-220 gpurun_out/bench_lib_${TAG}.json
for w in sd15-lcm4-512 sd15-lcm4-768; do
  python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${w}_${TAG}.json 2>/dev/null
  python -c "
import json,sys; d=json.loads(open(sys.argv[1]).readlines()[-1]); r=d['roofline']; print(sys.argv[2], round(d['value'],1), 'fps', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), ' igemm frac', round(r['frac'],3), 'step frac', round(r['step_frac'],3), ' lib', d.get('library_baseline'))" gpurun_out/bench_${w}_${TAG}.json $w
done
bash tools/ab.sh "" "B200SD_LANES=1" "B200SD_LANES=2" "B200SD_LANES=6" 2>&1 | tee gpurun_out/ab_${TAG}.txt
bash tools/ncu_frame.sh $TAG
bash tools/ncu_full.sh $TAG

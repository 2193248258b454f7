#!/bin/bash
# Round-end validation on the GPU box (run under gpurun): GPU parity suite, smoke, headline bench + reference / library arms.
# usage: bash tools/validate.sh [tag]   -> gpurun_out/bench_<tag>.json, gpurun_out/bench_ref_<tag>.json
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_{tag}.json").readlines()[-1]); r = d["roofline"]
print(f"{d['value']:.1f} fps  {d['ms_per_step']:.3f} ms ({d['frames_in_flight']} in flight; one at a time {d['sequential']['ms_per_frame']:.3f} ms)  "
      f"e2e {d['e2e']['value']:.1f} p50 {d['e2e']['p50_ms']:.2f} ms  igemm frac {r['frac']:.3f} (alone {r['one_frame_at_a_time']['frac']:.3f})  "
      f"library {d['library_baseline']['value']:.1f} fps  cpu {d['cpu_baseline']['value']:.3f} fps x{d['cpu_baseline']['cores']}  "
      f"clocks {d['clocks']['sm_mhz']}/{d['clocks']['sm_max_mhz']} {d['clocks']['reasons']}")
print(open(f"gpurun_out/bench_ref_{tag}.json").read()[:200])
PY

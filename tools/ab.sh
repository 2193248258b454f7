#!/bin/bash
# A/B the headline bench under different environment settings in ONE gpurun call (each variant is its own process: the
# tuning knobs are read once per process).  usage: bash tools/ab.sh "" "B2_NO_PDL=1" "B2_STAGE_KB=200" ...
for e in "$@"; do
  env $e python bench.py --steps ${AB_STEPS:-100} --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/ab_last.err | python -c "
import json, sys
d = json.loads(sys.stdin.readlines()[-1]); r = d['roofline']
o = r.get('other_kernels_in_graph', {})
print(repr(sys.argv[1]).ljust(44), f\"{d['value']:.1f} fps  {d['ms_per_step']:.3f} ms  seq {d.get('sequential',{}).get('ms_per_frame',0):.3f} ms  e2e {d['e2e']['value']:.1f} p50 {d['e2e']['p50_ms']:.2f}  igemm {r['kernel_ms_per_step']:.3f} ms ({r['kernel_launches_per_step']})  \" + '  '.join(f'{k} {v[\"ms_per_step\"]:.3f}' for k, v in o.items()))" "$e"
done

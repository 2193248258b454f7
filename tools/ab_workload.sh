#!/bin/bash
# same-box A/B of a non-headline workload: bash tools/ab_workload.sh <workload> "ENV=.." ...
W=$1; shift
for e in "$@"; do
  env $e python bench.py --workload $W --steps ${AB_STEPS:-60} --warmup 5 --no-cpu-baseline --no-library-baseline 2>gpurun_out/ab_last.err | python -c "
import json, sys
d = json.loads(sys.stdin.readlines()[-1]); r = d['roofline']; o = r.get('other_kernels_in_graph', {})
print(sys.argv[2], repr(sys.argv[1]).ljust(36), f\"{d['value']:.1f} fps  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.1f}  igemm {r['one_frame_at_a_time']['kernel_ms_per_step']:.3f} ms ({r['kernel_launches_per_step']})  \" + '  '.join(f'{k} {v[\"ms_per_step\"]:.3f}' for k, v in o.items()))" "$e" $W
done

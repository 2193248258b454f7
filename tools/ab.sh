#!/bin/bash
# A/B the headline bench under different environment settings in ONE gpurun call (each variant is its own process: the
# tuning knobs are read once per process).  usage: bash tools/ab.sh "" "B2_NO_PDL=1" "B2_STAGE_KB=200" ...
for e in "$@"; do
  env $e python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.readlines()[-1]); r = d['roofline']
print(repr(sys.argv[1]).ljust(28), f\"{d['value']:.1f} fps  {d['ms_per_step']:.3f} ms  igemm {r['kernel_ms_per_step']:.3f} ms\")" "$e"
done

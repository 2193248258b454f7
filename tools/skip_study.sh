#!/bin/bash
# marginal in-graph cost of each launch class: frame time with that class removed from the CUDA graph
for k in none groupnorm layernorm attn igemm smallconv upsample2x; do
  if [ $k = none ]; then unset B200SD_SKIP; else export B200SD_SKIP=$k; fi
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=$k', round(d['ms_per_step'],3), 'ms')"
done

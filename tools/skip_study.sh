#!/bin/bash
# marginal in-graph cost of each launch class: frame time with that class removed from the CUDA graph (timing only, output is garbage)
# usage: bash tools/skip_study.sh [kinds|names|all]
run() { python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['ms_per_step'],3), 'ms')" "$1"; }
what=${1:-all}
run "skip=none"
if [ $what != names ]; then for k in groupnorm layernorm attn igemm smallconv upsample2x; do B200SD_SKIP=$k run "skip kind=$k"; done; fi
if [ $what != kinds ]; then for n in "igemm vae." "attentions" "resnets" "geglu"; do B200SD_SKIP_NAME="$n" run "skip name=$n"; done; fi

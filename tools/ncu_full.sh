#!/bin/bash
# `ncu --set full` captures of representative launches (tools/ncu_ops.py), one GPU.  usage: bash tools/ncu_full.sh <tag> [ops...]
TAG=${1:-r02}; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'igemm_kernel|igemm_pair_kernel|tconv_kernel|attn_kernel|gn_cluster' -c 20 \
    -o gpurun_out/full_${TAG} -f python tools/ncu_ops.py "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log; ls -la gpurun_out/full_${TAG}.ncu-rep

#!/bin/bash
# One-off environment probe on the GPU box. Output -> gpurun_out/probe.txt
mkdir -p gpurun_out
{
  echo "== nvidia-smi"; nvidia-smi
  echo "== topo"; nvidia-smi topo -m
  echo "== codec libs"; ldconfig -p | grep -E 'nvcuvid|nvidia-encode|libcuda\.so' || echo "none"
  ls /usr/lib/x86_64-linux-gnu | grep -E 'nvcuvid|nvidia-encode' || true
  echo "== cpu"; nproc; lscpu | head -20; free -g | head -2
  echo "== torch"; python -c "import torch; print(torch.__version__, torch.cuda.is_available(), torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0)); print(torch.cuda.mem_get_info())"
} > gpurun_out/probe.txt 2>&1

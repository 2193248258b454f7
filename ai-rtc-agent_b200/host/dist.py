"""Multi-GPU plumbing: one process per GPU, independent video streams (stream i -> GPU i), weights
broadcast ONCE at init with NCCL over NVLink/NVSwitch, no per-frame collective.  The
reference has no multi-GPU path (dormant DataParallel hook, lib/wrapper.py:187-190)."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import arch as A
from . import weights as W


def env_rank() -> Tuple[int, int, int]:
    return int(os.getenv("RANK", "0")), int(os.getenv("WORLD_SIZE", "1")), int(os.getenv("LOCAL_RANK", "0"))


def init(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Join the job described by RANK / WORLD_SIZE / MASTER_* (torchrun).  No-op for a single process."""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def broadcast_state_dict(sd: Optional[Dict[str, torch.Tensor]], shapes: Dict[str, A.Shape], device: torch.device,
                         src: int = 0) -> Dict[str, torch.Tensor]:
    """All ranks return the same fp16 state dict living on `device`; only `src` needs to pass `sd`.
    One flat buffer, one broadcast (1.73 GB for the SD UNet: a few ms at NVLink rates)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    total = sum(int(torch.Size(s).numel()) for s in shapes.values())
    flat = torch.empty(total, dtype=torch.float16, device=device)
    if rank == src:
        if sd is None:
            raise ValueError("the source rank must provide the state dict")
        off = 0
        for k, shp in shapes.items():
            n = int(torch.Size(shp).numel())
            flat[off:off + n].copy_(sd[k].reshape(-1).to(torch.float16))
            off += n
    if world > 1:
        dist.broadcast(flat, src=src)
    out: Dict[str, torch.Tensor] = {}
    off = 0
    for k, shp in shapes.items():
        n = int(torch.Size(shp).numel())
        out[k] = flat[off:off + n].view(shp)
        off += n
    return out


def load_and_broadcast(model_id: str, device: torch.device, **resolve_kw):
    """Rank 0 resolves (or synthesises) the weights, every rank ends up with identical copies, and they are
    registered so that StreamDiffusionWrapper(model_id) on this rank uses them."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    arch = A.arch_for(model_id)
    unet_sd = vae_sd = None
    if rank == 0:
        arch, unet_sd, vae_sd, _ = W.resolve_weights(model_id, resolve_kw.get("vae_id"), resolve_kw.get("lcm_lora_id"),
                                                     resolve_kw.get("use_lcm_lora", True), resolve_kw.get("lora_dict"),
                                                     "turbo" in model_id)
    unet_sd = broadcast_state_dict(unet_sd, A.unet_param_shapes(arch), device)
    vae_sd = broadcast_state_dict(vae_sd, A.taesd_param_shapes(), device)
    W.register_preloaded(model_id, arch, unet_sd, vae_sd)
    return arch, unet_sd, vae_sd

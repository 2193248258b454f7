"""Seeded synthetic weights in diffusers' state-dict naming (no checkpoints are obtainable offline:
download.py:17-25 needs network).  Values are drawn in fp32 from a CPU generator and rounded to fp16,
so the oracle (fp32 math on the rounded values) and the CUDA path (fp16 storage) see identical
parameters.  Fan-in scaling keeps activations O(1) through ~200 layers.  Test/bench infrastructure."""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import taesd, unet


def _fill(shapes: Dict[str, tuple], seed: int, last_branch_gain: float) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        is_norm = ".norm" in name or name.startswith("conv_norm_out")
        if is_norm:
            t = (1.0 + 0.1 * torch.randn(shape, generator=g)) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            gain = 1.0
            # last layer of each residual branch: keep the residual stream tame
            if any(k in name for k in ("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "proj_out.weight",
                                       "conv.4.weight")):
                gain = last_branch_gain
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        sd[name] = t.to(torch.float16)
    return sd


def make_unet_weights(cfg: unet.UNetConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """fp16 state dict; use `to_float(sd)` for the oracle."""
    return _fill(unet.param_shapes(cfg), seed, 0.5)


def make_taesd_weights(seed: int = 4321) -> Dict[str, torch.Tensor]:
    sd = _fill(taesd.param_shapes(), seed, 0.7)
    # ReLU networks: He gain on the inner convs so the signal does not collapse
    g = torch.Generator().manual_seed(seed + 1)
    for name, t in list(sd.items()):
        if name.endswith("weight") and t.dim() == 4 and ("conv.0" in name or "conv.2" in name):
            fan_in = math.prod(t.shape[1:])
            sd[name] = (torch.randn(t.shape, generator=g) * math.sqrt(2.0 / fan_in)).to(torch.float16)
    # decoder head: centre the image around mid-grey so u8 outputs are not saturated
    last = max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))
    sd[f"decoder.layers.{last}.bias"] = torch.full((3,), 0.5, dtype=torch.float16)
    sd[f"decoder.layers.{last}.weight"] = (sd[f"decoder.layers.{last}.weight"].float() * 0.35).to(torch.float16)
    return sd


def to_float(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return {k: v.float() for k, v in sd.items()}


def make_prompt_embeds(dim: int, seed: int = 1, tokens: int = 77) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn((1, tokens, dim), generator=g).to(torch.float16)


def make_frame(height: int, width: int, seed: int = 0, smooth: bool = True) -> torch.Tensor:
    """(1,H,W,3) uint8 NHWC synthetic RGB frame: smooth colour field + noise (seeded)."""
    g = torch.Generator().manual_seed(seed)
    if not smooth:
        return torch.randint(0, 256, (1, height, width, 3), dtype=torch.uint8, generator=g)
    yy = torch.linspace(0, 1, height).view(height, 1, 1)
    xx = torch.linspace(0, 1, width).view(1, width, 1)
    ph = torch.rand(3, generator=g).view(1, 1, 3) * 6.28
    fr = (1.0 + 3.0 * torch.rand(3, generator=g)).view(1, 1, 3)
    img = 0.5 + 0.35 * torch.sin(6.28 * fr * (xx + 0.7 * yy) + ph) + 0.08 * torch.randn((height, width, 3), generator=g)
    return (img.clamp(0, 1) * 255).to(torch.uint8).unsqueeze(0)

"""fp32 restatement of diffusers 0.24 `AutoencoderTiny` (madebyollin/taesd; lib/wrapper.py:439-453,
699-707) -- models/autoencoder_tiny.py + models/vae.py (EncoderTiny / DecoderTiny /
AutoencoderTinyBlock).  SURVEY.md Appendix A.3.  Test infrastructure only (see oracle/__init__)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

CH = 64
ENC_BLOCKS = (1, 3, 3, 3)
DEC_BLOCKS = (3, 3, 3, 1)
LATENT = 4
SCALING_FACTOR = 1.0  # AutoencoderTiny.config.scaling_factor


def param_shapes() -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}

    def block(p):
        for k in (0, 2, 4):
            s[f"{p}.conv.{k}.weight"] = (CH, CH, 3, 3)
            s[f"{p}.conv.{k}.bias"] = (CH,)

    # encoder: nn.Sequential indices
    i = 0
    for stage, nblk in enumerate(ENC_BLOCKS):
        if stage == 0:
            s[f"encoder.layers.{i}.weight"] = (CH, 3, 3, 3); s[f"encoder.layers.{i}.bias"] = (CH,)
        else:
            s[f"encoder.layers.{i}.weight"] = (CH, CH, 3, 3)  # stride 2, bias=False
        i += 1
        for _ in range(nblk):
            block(f"encoder.layers.{i}")
            i += 1
    s[f"encoder.layers.{i}.weight"] = (LATENT, CH, 3, 3); s[f"encoder.layers.{i}.bias"] = (LATENT,)
    # decoder
    s["decoder.layers.0.weight"] = (CH, LATENT, 3, 3); s["decoder.layers.0.bias"] = (CH,)
    i = 2  # index 1 is the ReLU
    for stage, nblk in enumerate(DEC_BLOCKS):
        last = stage == len(DEC_BLOCKS) - 1
        for _ in range(nblk):
            block(f"decoder.layers.{i}")
            i += 1
        if not last:
            i += 1  # nn.Upsample
            s[f"decoder.layers.{i}.weight"] = (CH, CH, 3, 3)  # bias=False
        else:
            s[f"decoder.layers.{i}.weight"] = (3, CH, 3, 3); s[f"decoder.layers.{i}.bias"] = (3,)
        i += 1
    return s


def _block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderTinyBlock: relu(conv(relu(conv(relu(conv(x))))) + x) (skip = Identity, in == out)."""
    h = F.relu(F.conv2d(x, sd[f"{p}.conv.0.weight"], sd[f"{p}.conv.0.bias"], padding=1))
    h = F.relu(F.conv2d(h, sd[f"{p}.conv.2.weight"], sd[f"{p}.conv.2.bias"], padding=1))
    h = F.conv2d(h, sd[f"{p}.conv.4.weight"], sd[f"{p}.conv.4.bias"], padding=1)
    return F.relu(h + x)


def encode(sd: SD, images: torch.Tensor) -> torch.Tensor:
    """AutoencoderTiny.encode(x).latents: images in [-1,1] (B,3,H,W) -> latents (B,4,H/8,W/8).
    EncoderTiny.forward first maps x -> (x+1)/2."""
    x = (images + 1.0) / 2.0
    i = 0
    for stage, nblk in enumerate(ENC_BLOCKS):
        if stage == 0:
            x = F.conv2d(x, sd[f"encoder.layers.{i}.weight"], sd[f"encoder.layers.{i}.bias"], padding=1)
        else:
            x = F.conv2d(x, sd[f"encoder.layers.{i}.weight"], None, stride=2, padding=1)
        i += 1
        for _ in range(nblk):
            x = _block(sd, f"encoder.layers.{i}", x)
            i += 1
    return F.conv2d(x, sd[f"encoder.layers.{i}.weight"], sd[f"encoder.layers.{i}.bias"], padding=1)


def decode(sd: SD, latents: torch.Tensor) -> torch.Tensor:
    """AutoencoderTiny.decode(z).sample: DecoderTiny.forward = tanh(z/3)*3 -> layers -> 2y-1."""
    x = torch.tanh(latents / 3.0) * 3.0
    x = F.relu(F.conv2d(x, sd["decoder.layers.0.weight"], sd["decoder.layers.0.bias"], padding=1))
    i = 2
    for stage, nblk in enumerate(DEC_BLOCKS):
        last = stage == len(DEC_BLOCKS) - 1
        for _ in range(nblk):
            x = _block(sd, f"decoder.layers.{i}", x)
            i += 1
        if not last:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            i += 1
            x = F.conv2d(x, sd[f"decoder.layers.{i}.weight"], None, padding=1)
        else:
            x = F.conv2d(x, sd[f"decoder.layers.{i}.weight"], sd[f"decoder.layers.{i}.bias"], padding=1)
        i += 1
    return x * 2.0 - 1.0

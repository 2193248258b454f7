"""Architectures the reference can select (lib/wrapper.py:133: `"turbo" in model_id_or_path`) and the
parameter inventory of their checkpoints in diffusers' naming (used to validate loaded state dicts and
to build synthetic weights for benchmarking when no checkpoint is on disk -- download.py needs network).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Iterator, Tuple

import torch

Shape = Tuple[int, ...]


@dataclass(frozen=True)
class UNetArch:
    name: str
    block_out_channels: Tuple[int, int, int, int]
    heads: Tuple[int, int, int, int]          # diffusers `attention_head_dim` (= number of heads)
    cross_attention_dim: int
    use_linear_projection: bool
    down_attn: Tuple[int, int, int, int] = (1, 1, 1, 0)
    layers_per_block: int = 2
    norm_groups: int = 32
    ctx_tokens: int = 77


SD15 = UNetArch("sd15", (320, 640, 1280, 1280), (8, 8, 8, 8), 768, False)
SD_TURBO = UNetArch("sd-turbo", (320, 640, 1280, 1280), (5, 10, 20, 20), 1024, True)
# 1/5-width variants with the same topology (tests, smoke)
TINY_SD15 = UNetArch("tiny-sd15", (64, 128, 256, 256), (8, 8, 8, 8), 64, False)
TINY_TURBO = UNetArch("tiny-turbo", (64, 128, 256, 256), (1, 2, 4, 4), 128, True)


def arch_for(model_id: str) -> UNetArch:
    if model_id.rstrip("/").rsplit("/", 1)[-1].startswith("tiny"):   # a model id or a local directory name
        return TINY_TURBO if "turbo" in model_id else TINY_SD15
    return SD_TURBO if "turbo" in model_id else SD15


def _norm(prefix: str, c: int) -> Iterator[Tuple[str, Shape]]:
    yield prefix + ".weight", (c,)
    yield prefix + ".bias", (c,)


def _conv(prefix: str, co: int, ci: int, k: int, bias: bool = True) -> Iterator[Tuple[str, Shape]]:
    yield prefix + ".weight", (co, ci, k, k)
    if bias:
        yield prefix + ".bias", (co,)


def _linear(prefix: str, co: int, ci: int, bias: bool = True) -> Iterator[Tuple[str, Shape]]:
    yield prefix + ".weight", (co, ci)
    if bias:
        yield prefix + ".bias", (co,)


def _resnet(prefix: str, ci: int, co: int, tdim: int) -> Iterator[Tuple[str, Shape]]:
    yield from _norm(prefix + ".norm1", ci)
    yield from _conv(prefix + ".conv1", co, ci, 3)
    yield from _linear(prefix + ".time_emb_proj", co, tdim)
    yield from _norm(prefix + ".norm2", co)
    yield from _conv(prefix + ".conv2", co, co, 3)
    if ci != co:
        yield from _conv(prefix + ".conv_shortcut", co, ci, 1)


def _transformer(prefix: str, c: int, a: UNetArch) -> Iterator[Tuple[str, Shape]]:
    yield from _norm(prefix + ".norm", c)
    for proj in ("proj_in", "proj_out"):
        if a.use_linear_projection:
            yield from _linear(f"{prefix}.{proj}", c, c)
        else:
            yield from _conv(f"{prefix}.{proj}", c, c, 1)
    blk = prefix + ".transformer_blocks.0"
    for i in (1, 2, 3):
        yield from _norm(f"{blk}.norm{i}", c)
    for attn, kv in (("attn1", c), ("attn2", a.cross_attention_dim)):
        yield from _linear(f"{blk}.{attn}.to_q", c, c, bias=False)
        yield from _linear(f"{blk}.{attn}.to_k", c, kv, bias=False)
        yield from _linear(f"{blk}.{attn}.to_v", c, kv, bias=False)
        yield from _linear(f"{blk}.{attn}.to_out.0", c, c)
    yield from _linear(f"{blk}.ff.net.0.proj", 8 * c, c)
    yield from _linear(f"{blk}.ff.net.2", c, 4 * c)


def unet_param_shapes(a: UNetArch) -> Dict[str, Shape]:
    ch = a.block_out_channels
    tdim = 4 * ch[0]
    out: Dict[str, Shape] = {}
    out.update(_conv("conv_in", ch[0], 4, 3))
    out.update(_linear("time_embedding.linear_1", tdim, ch[0]))
    out.update(_linear("time_embedding.linear_2", tdim, tdim))
    skip = [ch[0]]
    cur = ch[0]
    for i, co in enumerate(ch):
        for j in range(a.layers_per_block):
            out.update(_resnet(f"down_blocks.{i}.resnets.{j}", cur, co, tdim))
            cur = co
            if a.down_attn[i]:
                out.update(_transformer(f"down_blocks.{i}.attentions.{j}", co, a))
            skip.append(cur)
        if i < len(ch) - 1:
            out.update(_conv(f"down_blocks.{i}.downsamplers.0.conv", co, co, 3))
            skip.append(cur)
    out.update(_resnet("mid_block.resnets.0", cur, cur, tdim))
    out.update(_transformer("mid_block.attentions.0", cur, a))
    out.update(_resnet("mid_block.resnets.1", cur, cur, tdim))
    for i, co in enumerate(reversed(ch)):
        for j in range(a.layers_per_block + 1):
            out.update(_resnet(f"up_blocks.{i}.resnets.{j}", cur + skip.pop(), co, tdim))
            cur = co
            if a.down_attn[len(ch) - 1 - i]:
                out.update(_transformer(f"up_blocks.{i}.attentions.{j}", co, a))
        if i < len(ch) - 1:
            out.update(_conv(f"up_blocks.{i}.upsamplers.0.conv", co, co, 3))
    out.update(_norm("conv_norm_out", ch[0]))
    out.update(_conv("conv_out", 4, ch[0], 3))
    return out


def taesd_param_shapes() -> Dict[str, Shape]:
    """madebyollin/taesd AutoencoderTiny: nn.Sequential indices as diffusers builds them."""
    out: Dict[str, Shape] = {}

    def block(p: str):
        for k in (0, 2, 4):
            out.update(_conv(f"{p}.conv.{k}", 64, 64, 3))

    idx = 0
    for stage, nblk in enumerate((1, 3, 3, 3)):
        out.update(_conv(f"encoder.layers.{idx}", 64, 3 if stage == 0 else 64, 3, bias=(stage == 0)))
        idx += 1
        for _ in range(nblk):
            block(f"encoder.layers.{idx}")
            idx += 1
    out.update(_conv(f"encoder.layers.{idx}", 4, 64, 3))
    out.update(_conv("decoder.layers.0", 64, 4, 3))
    idx = 2
    for stage, nblk in enumerate((3, 3, 3, 1)):
        for _ in range(nblk):
            block(f"decoder.layers.{idx}")
            idx += 1
        if stage < 3:
            idx += 1  # nn.Upsample
            out.update(_conv(f"decoder.layers.{idx}", 64, 64, 3, bias=False))
        else:
            out.update(_conv(f"decoder.layers.{idx}", 3, 64, 3))
        idx += 1
    return out


def synthetic_state_dict(shapes: Dict[str, Shape], seed: int, relu_net: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded fp16 weights with fan-in scaling (activations stay O(1) through the whole network); used by
    bench.py / smoke when no checkpoint exists.  Not a trained model: outputs are noise-like images."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        leaf = name.rsplit(".", 1)[-1]
        if ".norm" in name or name.startswith("conv_norm_out"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            gain = math.sqrt(2.0) if relu_net else 1.0
            if any(k in name for k in ("conv2.", "to_out.0.", "ff.net.2.", "proj_out.", "conv.4.")):
                gain = 0.5 if not relu_net else 0.7
            t = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        sd[name] = t.to(torch.float16)
    if relu_net:  # keep the decoded image around mid-grey
        last = max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.layers."))
        sd[f"decoder.layers.{last}.bias"] = torch.full((3,), 0.5, dtype=torch.float16)
        sd[f"decoder.layers.{last}.weight"] = (sd[f"decoder.layers.{last}.weight"].float() * 0.25).to(torch.float16)
    return sd


def validate_state_dict(sd: Dict[str, torch.Tensor], shapes: Dict[str, Shape], what: str) -> None:
    missing = [k for k in shapes if k not in sd]
    if missing:
        raise ValueError(f"{what}: {len(missing)} missing tensors, e.g. {missing[:4]}")
    for k, shp in shapes.items():
        if tuple(sd[k].shape) != tuple(shp):
            # diffusers stores 1x1 conv projections as (c, c, 1, 1); linear ones as (c, c)
            if math.prod(sd[k].shape) != math.prod(shp):
                raise ValueError(f"{what}: {k} has shape {tuple(sd[k].shape)}, expected {shp}")

"""fp32 restatement of diffusers 0.24 `UNet2DConditionModel` for the two architectures the reference
runs:

  * SD-1.5 family (lykon/dreamshaper-8, agent.py:443): cross_attention_dim 768, 8 heads everywhere,
    1x1-conv proj_in/proj_out;
  * SD-Turbo / SD-2.1-base (model id containing "turbo", lib/wrapper.py:133): cross_attention_dim 1024,
    head dim 64 (5/10/20/20 heads), Linear proj_in/proj_out.

Functional style over a flat state dict that uses diffusers' key names, so a real checkpoint's
`unet/diffusion_pytorch_model.safetensors` loads unchanged.  Files restated (not present under
/root/reference): models/unet_2d_condition.py, unet_2d_blocks.py, resnet.py, transformer_2d.py,
attention.py, attention_processor.py, embeddings.py.  Test infrastructure only (see oracle/__init__).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
FUSED_ATTENTION = False   # True: F.scaled_dot_product_attention instead of the explicit softmax(QK^T)V (same math, library kernels)


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    down_attn: Tuple[bool, ...] = (True, True, True, False)   # CrossAttnDownBlock2D x3 + DownBlock2D
    heads: Tuple[int, ...] = (8, 8, 8, 8)                      # `attention_head_dim` == number of heads
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    norm_groups: int = 32
    norm_eps: float = 1e-5
    time_dim_mult: int = 4

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * self.time_dim_mult

    @property
    def up_attn(self) -> Tuple[bool, ...]:
        return tuple(reversed(self.down_attn))


SD15 = UNetConfig()
SD_TURBO = UNetConfig(heads=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True)


def tiny_config(turbo: bool) -> UNetConfig:
    """Same topology at 1/5 width (64-channel granularity kept) for fast tests."""
    if turbo:
        return UNetConfig(block_out_channels=(64, 128, 256, 256), heads=(1, 2, 4, 4),
                          cross_attention_dim=128, use_linear_projection=True)
    return UNetConfig(block_out_channels=(64, 128, 256, 256), heads=(8, 8, 8, 8), cross_attention_dim=64,
                      use_linear_projection=False)


def config_for(model_id: str) -> UNetConfig:
    """lib/wrapper.py:133 -- `"turbo" in model_id_or_path` selects the SD-Turbo behaviour."""
    return SD_TURBO if "turbo" in model_id else SD15


# ------------------------------------------------------------------------------------------------
# parameter inventory (name -> shape), in diffusers' naming
def _resnet_shapes(p: str, cin: int, cout: int, temb: int, out: Dict[str, Tuple[int, ...]]):
    out[p + "norm1.weight"] = (cin,); out[p + "norm1.bias"] = (cin,)
    out[p + "conv1.weight"] = (cout, cin, 3, 3); out[p + "conv1.bias"] = (cout,)
    out[p + "time_emb_proj.weight"] = (cout, temb); out[p + "time_emb_proj.bias"] = (cout,)
    out[p + "norm2.weight"] = (cout,); out[p + "norm2.bias"] = (cout,)
    out[p + "conv2.weight"] = (cout, cout, 3, 3); out[p + "conv2.bias"] = (cout,)
    if cin != cout:
        out[p + "conv_shortcut.weight"] = (cout, cin, 1, 1); out[p + "conv_shortcut.bias"] = (cout,)


def _attn_shapes(p: str, c: int, cfg: UNetConfig, out: Dict[str, Tuple[int, ...]]):
    out[p + "norm.weight"] = (c,); out[p + "norm.bias"] = (c,)
    if cfg.use_linear_projection:
        out[p + "proj_in.weight"] = (c, c); out[p + "proj_out.weight"] = (c, c)
    else:
        out[p + "proj_in.weight"] = (c, c, 1, 1); out[p + "proj_out.weight"] = (c, c, 1, 1)
    out[p + "proj_in.bias"] = (c,); out[p + "proj_out.bias"] = (c,)
    t = p + "transformer_blocks.0."
    for n in ("norm1", "norm2", "norm3"):
        out[t + n + ".weight"] = (c,); out[t + n + ".bias"] = (c,)
    for a, kv in (("attn1", c), ("attn2", cfg.cross_attention_dim)):
        out[t + a + ".to_q.weight"] = (c, c)
        out[t + a + ".to_k.weight"] = (c, kv)
        out[t + a + ".to_v.weight"] = (c, kv)
        out[t + a + ".to_out.0.weight"] = (c, c); out[t + a + ".to_out.0.bias"] = (c,)
    out[t + "ff.net.0.proj.weight"] = (8 * c, c); out[t + "ff.net.0.proj.bias"] = (8 * c,)
    out[t + "ff.net.2.weight"] = (c, 4 * c); out[t + "ff.net.2.bias"] = (c,)


def param_shapes(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    s: Dict[str, Tuple[int, ...]] = {}
    ch = cfg.block_out_channels
    temb = cfg.time_embed_dim
    s["conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3); s["conv_in.bias"] = (ch[0],)
    s["time_embedding.linear_1.weight"] = (temb, ch[0]); s["time_embedding.linear_1.bias"] = (temb,)
    s["time_embedding.linear_2.weight"] = (temb, temb); s["time_embedding.linear_2.bias"] = (temb,)
    # down
    skip_ch: List[int] = [ch[0]]
    cur = ch[0]
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            _resnet_shapes(f"down_blocks.{i}.resnets.{j}.", cur, co, temb, s)
            cur = co
            if cfg.down_attn[i]:
                _attn_shapes(f"down_blocks.{i}.attentions.{j}.", co, cfg, s)
            skip_ch.append(cur)
        if i != len(ch) - 1:
            s[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (co, co, 3, 3)
            s[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (co,)
            skip_ch.append(cur)
    # mid
    _resnet_shapes("mid_block.resnets.0.", cur, cur, temb, s)
    _attn_shapes("mid_block.attentions.0.", cur, cfg, s)
    _resnet_shapes("mid_block.resnets.1.", cur, cur, temb, s)
    # up
    rev = list(reversed(ch))
    for i, co in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            sk = skip_ch.pop()
            _resnet_shapes(f"up_blocks.{i}.resnets.{j}.", cur + sk, co, temb, s)
            cur = co
            if cfg.up_attn[i]:
                _attn_shapes(f"up_blocks.{i}.attentions.{j}.", co, cfg, s)
        if i != len(ch) - 1:
            s[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (co, co, 3, 3)
            s[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (co,)
    s["conv_norm_out.weight"] = (ch[0],); s["conv_norm_out.bias"] = (ch[0],)
    s["conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3); s["conv_out.bias"] = (cfg.out_channels,)
    return s


def param_count(cfg: UNetConfig) -> int:
    return sum(math.prod(v) for v in param_shapes(cfg).values())


# ------------------------------------------------------------------------------------------------
# forward
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    """embeddings.py get_timestep_embedding with flip_sin_to_cos=True, freq_shift=0: [cos | sin]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t.float()[:, None] * freqs[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def time_embed(sd: SD, cfg: UNetConfig, t: torch.Tensor) -> torch.Tensor:
    e = timestep_embedding(t, cfg.block_out_channels[0]).to(sd["time_embedding.linear_1.weight"].dtype)  # diffusers: t_emb.to(sample.dtype)
    e = F.linear(e, sd["time_embedding.linear_1.weight"], sd["time_embedding.linear_1.bias"])
    return F.linear(F.silu(e), sd["time_embedding.linear_2.weight"], sd["time_embedding.linear_2.bias"])


def resnet(sd: SD, p: str, cfg: UNetConfig, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    h = F.silu(F.group_norm(x, cfg.norm_groups, sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.norm_eps))
    h = F.conv2d(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1)
    te = F.linear(F.silu(emb), sd[p + "time_emb_proj.weight"], sd[p + "time_emb_proj.bias"])
    h = h + te[:, :, None, None]
    h = F.silu(F.group_norm(h, cfg.norm_groups, sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.norm_eps))
    h = F.conv2d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], padding=1)
    if (p + "conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + "conv_shortcut.weight"], sd[p + "conv_shortcut.bias"])
    return x + h


def attention(sd: SD, p: str, heads: int, x: torch.Tensor, ctx: torch.Tensor) -> torch.Tensor:
    """attention_processor.Attention: bias-free q/k/v, softmax(q k^T / sqrt(d)) v, to_out.0 with bias."""
    b, n, c = x.shape
    d = c // heads
    q = F.linear(x, sd[p + "to_q.weight"]).view(b, n, heads, d).transpose(1, 2)
    k = F.linear(ctx, sd[p + "to_k.weight"]).view(b, -1, heads, d).transpose(1, 2)
    v = F.linear(ctx, sd[p + "to_v.weight"]).view(b, -1, heads, d).transpose(1, 2)
    if FUSED_ATTENTION:   # library leg (torch_gpu.py): flash / memory-efficient SDPA, what diffusers' AttnProcessor2_0 calls
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
    else:
        s = torch.softmax((q @ k.transpose(-1, -2)) * (d ** -0.5), dim=-1)
        o = (s @ v).transpose(1, 2).reshape(b, n, c)
    return F.linear(o, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def transformer(sd: SD, p: str, cfg: UNetConfig, heads: int, x: torch.Tensor, ctx: torch.Tensor) -> torch.Tensor:
    """Transformer2DModel with one BasicTransformerBlock (self-attn, cross-attn, GEGLU feed-forward)."""
    b, c, hh, ww = x.shape
    res = x
    h = F.group_norm(x, cfg.norm_groups, sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-6)
    if cfg.use_linear_projection:
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        h = F.linear(h, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
    else:
        h = F.conv2d(h, sd[p + "proj_in.weight"], sd[p + "proj_in.bias"])
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    t = p + "transformer_blocks.0."
    n1 = F.layer_norm(h, (c,), sd[t + "norm1.weight"], sd[t + "norm1.bias"], 1e-5)
    h = h + attention(sd, t + "attn1.", heads, n1, n1)
    n2 = F.layer_norm(h, (c,), sd[t + "norm2.weight"], sd[t + "norm2.bias"], 1e-5)
    h = h + attention(sd, t + "attn2.", heads, n2, ctx)
    n3 = F.layer_norm(h, (c,), sd[t + "norm3.weight"], sd[t + "norm3.bias"], 1e-5)
    proj = F.linear(n3, sd[t + "ff.net.0.proj.weight"], sd[t + "ff.net.0.proj.bias"])
    val, gate = proj.chunk(2, dim=-1)
    h = h + F.linear(val * F.gelu(gate), sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"])
    if cfg.use_linear_projection:
        h = F.linear(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
        h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    else:
        h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        h = F.conv2d(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return h + res


def unet_forward(sd: SD, cfg: UNetConfig, sample: torch.Tensor, timesteps: torch.Tensor,
                 encoder_hidden_states: torch.Tensor, taps: Dict[str, torch.Tensor] | None = None) -> torch.Tensor:
    """eps = UNet(sample (B,4,h,w), timesteps (B,), encoder_hidden_states (B,77,D)).  `taps`, if given,
    collects named intermediate activations (NCHW) for layer-by-layer parity debugging."""
    ch = cfg.block_out_channels
    emb = time_embed(sd, cfg, timesteps)
    ctx = encoder_hidden_states
    h = F.conv2d(sample, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    skips = [h]
    if taps is not None:
        taps["conv_in"] = h
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            h = resnet(sd, f"down_blocks.{i}.resnets.{j}.", cfg, h, emb)
            if cfg.down_attn[i]:
                h = transformer(sd, f"down_blocks.{i}.attentions.{j}.", cfg, cfg.heads[i], h, ctx)
            skips.append(h)
            if taps is not None:
                taps[f"down.{i}.{j}"] = h
        if i != len(ch) - 1:
            p = f"down_blocks.{i}.downsamplers.0.conv."
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], stride=2, padding=1)
            skips.append(h)
    h = resnet(sd, "mid_block.resnets.0.", cfg, h, emb)
    h = transformer(sd, "mid_block.attentions.0.", cfg, cfg.heads[-1], h, ctx)
    h = resnet(sd, "mid_block.resnets.1.", cfg, h, emb)
    if taps is not None:
        taps["mid"] = h
    rheads = list(reversed(cfg.heads))
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(sd, f"up_blocks.{i}.resnets.{j}.", cfg, h, emb)
            if cfg.up_attn[i]:
                h = transformer(sd, f"up_blocks.{i}.attentions.{j}.", cfg, rheads[i], h, ctx)
            if taps is not None:
                taps[f"up.{i}.{j}"] = h
        if i != len(ch) - 1:
            p = f"up_blocks.{i}.upsamplers.0.conv."
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[p + "weight"], sd[p + "bias"], padding=1)
    h = F.silu(F.group_norm(h, cfg.norm_groups, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], cfg.norm_eps))
    return F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)

"""CPU tests of the oracle (no GPU): everything that CAN be pinned without the reference's un-installable
dependencies -- published architecture facts, closed-form scheduler constants, independent torch.nn
re-derivations of each block, the stream-batch semantics, and the committed golden fixtures."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import pipeline as opipe
from oracle import stream as ostream
from oracle import taesd as otaesd
from oracle import unet as ounet
from oracle import weights as ow

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- published facts ------------------------------------------------------------------------------------
def test_parameter_counts_match_published_models():
    # runwayml/stable-diffusion-v1-5 UNet: 859,520,964 parameters; SD-2.1-base / SD-Turbo UNet: 865,910,724
    assert ounet.param_count(ounet.SD15) == 859_520_964
    assert ounet.param_count(ounet.SD_TURBO) == 865_910_724
    # madebyollin/taesd: encoder + decoder = 2,445,063 parameters
    assert sum(math.prod(s) for s in otaesd.param_shapes().values()) == 2_445_063


def test_layer_counts():
    keys = ounet.param_shapes(ounet.SD15).keys()
    assert len({k.split(".conv1.")[0] for k in keys if ".conv1.weight" in k}) == 22          # resnets
    assert len({k.split(".transformer_blocks")[0] for k in keys if "transformer_blocks" in k}) == 16
    assert sum(1 for k in keys if "downsamplers" in k and k.endswith("weight")) == 3
    assert sum(1 for k in keys if "upsamplers" in k and k.endswith("weight")) == 3


def test_lcm_timestep_table():
    ts = ostream.lcm_timesteps(50)
    assert ts == [999 - 20 * i for i in range(50)]
    assert [ts[i] for i in (18, 26, 35, 45)] == [639, 479, 299, 99]   # lib/pipeline.py:12 defaults


def test_scheduler_constants_closed_form():
    ac = ostream.alphas_cumprod()
    assert abs(ac[0].item() - (1 - 0.00085)) < 1e-7
    assert abs(ac[999].item() - 0.0046601) < 2e-6          # SD scaled-linear schedule, final alpha_bar
    assert torch.all(ac[1:] < ac[:-1])
    c_skip, c_out = ostream.boundary_scalings(99)
    assert abs(c_skip - 0.25 / (990.0 ** 2 + 0.25)) < 1e-12
    assert abs(c_out - 990.0 / math.sqrt(990.0 ** 2 + 0.25)) < 1e-12


def test_timestep_embedding_layout():
    e = ounet.timestep_embedding(torch.tensor([0.0, 639.0]), 320)
    assert torch.allclose(e[0, :160], torch.ones(160)) and torch.allclose(e[0, 160:], torch.zeros(160))  # [cos | sin]
    k = 37
    f = math.exp(-math.log(10000.0) * k / 160)
    assert abs(e[1, k].item() - math.cos(639 * f)) < 1e-5 and abs(e[1, 160 + k].item() - math.sin(639 * f)) < 1e-5


# ---- independent re-derivations with torch.nn modules ---------------------------------------------------------
def _sd_for(cfg, seed=0):
    return ow.to_float(ow.make_unet_weights(cfg, seed))


def test_resnet_block_vs_nn_modules():
    cfg = ounet.tiny_config(True)
    sd = _sd_for(cfg)
    p = "down_blocks.1.resnets.0."      # 64 -> 128 with 1x1 shortcut
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 64, 8, 8, generator=g)
    emb = torch.randn(2, cfg.time_embed_dim, generator=g)
    n1, n2 = nn.GroupNorm(32, 64, eps=1e-5), nn.GroupNorm(32, 128, eps=1e-5)
    c1, c2, sc = nn.Conv2d(64, 128, 3, padding=1), nn.Conv2d(128, 128, 3, padding=1), nn.Conv2d(64, 128, 1)
    tp = nn.Linear(cfg.time_embed_dim, 128)
    for mod, name in ((n1, "norm1"), (n2, "norm2"), (c1, "conv1"), (c2, "conv2"), (sc, "conv_shortcut"), (tp, "time_emb_proj")):
        mod.weight.data = sd[p + name + ".weight"].clone()
        mod.bias.data = sd[p + name + ".bias"].clone()
    h = c1(F.silu(n1(x))) + tp(F.silu(emb))[:, :, None, None]
    ref = sc(x) + c2(F.silu(n2(h)))
    assert torch.allclose(ounet.resnet(sd, p, cfg, x, emb), ref, atol=1e-5)


def test_attention_vs_nn_multihead_attention():
    cfg = ounet.tiny_config(True)
    sd = _sd_for(cfg)
    p = "down_blocks.1.attentions.0.transformer_blocks.0.attn1."
    c, heads = 128, 2
    mha = nn.MultiheadAttention(c, heads, bias=True, batch_first=True)
    mha.in_proj_weight.data = torch.cat([sd[p + "to_q.weight"], sd[p + "to_k.weight"], sd[p + "to_v.weight"]])
    mha.in_proj_bias.data.zero_()
    mha.out_proj.weight.data = sd[p + "to_out.0.weight"].clone()
    mha.out_proj.bias.data = sd[p + "to_out.0.bias"].clone()
    x = torch.randn(2, 64, c, generator=torch.Generator().manual_seed(2))
    ref, _ = mha(x, x, x, need_weights=False)
    assert torch.allclose(ounet.attention(sd, p, heads, x, x), ref, atol=2e-5)


def test_geglu_feed_forward_uses_erf_gelu():
    cfg = ounet.tiny_config(True)
    sd = _sd_for(cfg)
    t = "mid_block.attentions.0.transformer_blocks.0."
    x = torch.randn(1, 4, 256, generator=torch.Generator().manual_seed(3))
    proj = F.linear(x, sd[t + "ff.net.0.proj.weight"], sd[t + "ff.net.0.proj.bias"])
    a, gate = proj[..., :1024], proj[..., 1024:]
    erf_gelu = 0.5 * gate * (1 + torch.erf(gate / math.sqrt(2)))
    ref = F.linear(a * erf_gelu, sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"])
    val, g2 = proj.chunk(2, dim=-1)
    got = F.linear(val * F.gelu(g2), sd[t + "ff.net.2.weight"], sd[t + "ff.net.2.bias"])
    assert torch.allclose(got, ref, atol=1e-5)


def test_taesd_sequential_equivalence():
    """Build the nn.Sequential exactly as diffusers' EncoderTiny / DecoderTiny do and load the oracle's state dict."""
    sd = ow.to_float(ow.make_taesd_weights())

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Sequential(nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(), nn.Conv2d(64, 64, 3, padding=1), nn.ReLU(),
                                      nn.Conv2d(64, 64, 3, padding=1))
            self.skip = nn.Identity()
            self.fuse = nn.ReLU()

        def forward(self, x):
            return self.fuse(self.conv(x) + self.skip(x))

    enc_layers = []
    for i, nb in enumerate((1, 3, 3, 3)):
        enc_layers.append(nn.Conv2d(3, 64, 3, padding=1) if i == 0 else nn.Conv2d(64, 64, 3, padding=1, stride=2, bias=False))
        enc_layers += [Block() for _ in range(nb)]
    enc_layers.append(nn.Conv2d(64, 4, 3, padding=1))
    enc = nn.Sequential(*enc_layers)
    dec_layers = [nn.Conv2d(4, 64, 3, padding=1), nn.ReLU()]
    for i, nb in enumerate((3, 3, 3, 1)):
        dec_layers += [Block() for _ in range(nb)]
        last = i == 3
        if not last:
            dec_layers.append(nn.Upsample(scale_factor=2))
        dec_layers.append(nn.Conv2d(64, 3 if last else 64, 3, padding=1, bias=last))
    dec = nn.Sequential(*dec_layers)
    holder = nn.Module()
    holder.encoder, holder.decoder = nn.Module(), nn.Module()
    holder.encoder.layers, holder.decoder.layers = enc, dec
    missing, unexpected = holder.load_state_dict(sd, strict=True), None
    x = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(4)) * 2 - 1
    z = enc((x + 1) / 2)
    assert torch.allclose(otaesd.encode(sd, x), z, atol=1e-5)
    y = dec(torch.tanh(z / 3) * 3) * 2 - 1
    assert torch.allclose(otaesd.decode(sd, z), y, atol=1e-5)
    assert z.shape == (1, 4, 4, 4) and y.shape == (1, 3, 32, 32)


def test_unet_shapes_and_batch_independence():
    cfg = ounet.tiny_config(False)
    sd = _sd_for(cfg)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 4, 8, 8, generator=g)
    ctx = torch.randn(1, 77, cfg.cross_attention_dim, generator=g).repeat(3, 1, 1)
    t = torch.tensor([639, 479, 299])
    out = ounet.unet_forward(sd, cfg, x, t, ctx)
    assert out.shape == x.shape
    one = ounet.unet_forward(sd, cfg, x[1:2], t[1:2], ctx[1:2])
    assert torch.allclose(out[1:2], one, atol=1e-4)  # stream-batch slots do not interact inside the UNet


# ---- StreamDiffusion loop semantics ----------------------------------------------------------------------
def _tiny_stream(t_index_list, hw=64):
    cfg = ounet.tiny_config(True)
    orc = ostream.StreamOracle(_sd_for(cfg), cfg, ow.to_float(ow.make_taesd_weights()), t_index_list, hw, hw)
    orc.prepare(ow.make_prompt_embeds(cfg.cross_attention_dim).float(), guidance_scale=0.0)
    return orc


def test_stream_batch_output_lag():
    """With T denoising slots the image of input frame n leaves at call n+T-1 (SURVEY a-6)."""
    T = 3
    frames = [ow.make_frame(64, 64, seed=i) for i in range(6)]
    alt = ow.make_frame(64, 64, seed=99)
    a, b = _tiny_stream([10, 25, 40]), _tiny_stream([10, 25, 40])
    outs_a = [opipe.frame_to_u8(a, f) for f in frames]
    frames_b = list(frames)
    frames_b[1] = alt                     # change only frame 1
    outs_b = [opipe.frame_to_u8(b, f) for f in frames_b]
    differs = [not torch.equal(x, y) for x, y in zip(outs_a, outs_b)]
    assert differs == [False, False, False, True, False, False]   # only call 1 + (T-1) = 3 sees it


def test_prepare_state_and_update_quirk():
    orc = _tiny_stream([18, 26, 35, 45])
    assert orc.x_t_latent_buffer.shape == (3, 4, 8, 8) and float(orc.x_t_latent_buffer.abs().max()) == 0.0
    assert orc.sub_timesteps == [639, 479, 299, 99]
    a0 = orc.alpha_prod_t_sqrt.clone()
    orc.update_t_index_list([0, 10, 20, 30])
    assert orc.sub_timesteps == [999, 799, 599, 399]
    assert torch.equal(orc.alpha_prod_t_sqrt, a0)        # lib/wrapper.py:389-407 leaves the scalars untouched
    g = torch.Generator().manual_seed(2)
    assert torch.equal(orc.init_noise, torch.randn((4, 4, 8, 8), generator=g))  # seed 2, CPU generator


def test_pre_post_semantics():
    frame = torch.tensor([[[[0, 128, 255]]]], dtype=torch.uint8)       # (1,1,1,3) NHWC
    x = opipe.preprocess(frame)
    assert x.shape == (3, 1, 1) and x.dtype == torch.float32
    assert torch.allclose(x.flatten(), torch.tensor([0.0, 128 / 255, 1.0]))
    y = torch.tensor([0.0, 0.5, 0.999, 1.0, 1.7, -0.2]).view(6, 1, 1)
    u8 = opipe.postprocess(y)
    assert u8.shape == (1, 6, 1, 1)
    assert u8.flatten().tolist() == [0, 127, 254, 255, 255, 0]         # truncation, not rounding
    img = torch.tensor([[[[-1.0]], [[0.0]], [[3.0]]]])                 # (1,3,1,1) in [-1,1]
    assert opipe.denormalize_pt(img).flatten().tolist() == [0.0, 0.5, 1.0]


def test_image_preprocess_resize_and_skip_normalise():
    x = torch.rand(3, 10, 12)
    y = ostream.image_preprocess(x, 20, 24)
    assert y.shape == (1, 3, 20, 24) and float(y.min()) >= -1 and float(y.max()) <= 1
    assert torch.equal(y[0, :, ::2, ::2], 2 * x - 1)                    # nearest neighbour
    neg = torch.rand(1, 3, 4, 4) - 0.5
    assert torch.equal(ostream.image_preprocess(neg, 4, 4), neg)       # already signed: left untouched


# ---- committed golden vectors ------------------------------------------------------------------------------
@pytest.mark.parametrize("name,turbo,tl", [("tiny_turbo_T1", True, [32]), ("tiny_sd15_T4", False, [18, 26, 35, 45])])
def test_oracle_reproduces_golden(name, turbo, tl):
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    cfg = ounet.tiny_config(turbo)
    orc = ostream.StreamOracle(_sd_for(cfg, 1234), cfg, ow.to_float(ow.make_taesd_weights()), tl, 128, 128)
    orc.prepare(ow.make_prompt_embeds(cfg.cross_attention_dim).float(), guidance_scale=0.0, seed=2)
    orc.init_noise = orc.init_noise.half().float()
    assert list(gold["sub_timesteps"]) == orc.sub_timesteps
    for i in range(3):
        u8 = opipe.frame_to_u8(orc, ow.make_frame(128, 128, seed=i)).numpy()
        d = np.abs(u8.astype(np.int32) - gold["u8"][i:i + 1].astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 1e-3, f"{name} frame {i}: max {d.max()}"   # fp32 summation-order slack
        assert np.allclose(orc.last["eps"].numpy(), gold["eps"][i], atol=2e-4)

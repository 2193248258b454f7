"""GPU parity of the attention and SIMT helper kernels through the C ABI against PyTorch fp32 ops.

Tolerances: fp16 operands, fp32 math; outputs rounded to fp16 once.  Attention additionally rounds
P to fp16 before P.V (as every fp16 flash-attention does): abs 2e-3 on O(1) outputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _ops():
    from ai_rtc_agent_b200.host import ops
    return ops


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


@pytest.mark.parametrize("nb,heads,seq,d,dp", [
    (1, 1, 128, 64, 64),     # one q tile, one kv block
    (1, 2, 256, 64, 64),     # two kv blocks: online-softmax rescale, S double buffer
    (1, 5, 4096, 64, 64),    # SD-Turbo 64x64 latent self-attention
    (2, 10, 1024, 64, 64),   # batch 2
    (1, 20, 64, 64, 64),     # 8x8 level: half-empty q tile, masked kv tail
    (1, 4, 576, 64, 64),     # 768-class odd length (tail masking)
    (1, 8, 256, 40, 64),     # SD-1.5 head dim 40 zero-padded to 64
    (1, 8, 256, 80, 128),    # SD-1.5 head dim 80 -> 128
    (2, 8, 384, 160, 192),   # SD-1.5 head dim 160 -> 192 (BKV 64 variant)
])
def test_self_attention(cuda, nb, heads, seq, d, dp):
    ops = _ops()
    q = _rand((nb, heads, seq, d), cuda, 1).half()
    k = _rand((nb, heads, seq, d), cuda, 2).half()
    v = _rand((nb, heads, seq, d), cuda, 3).half()
    # kernel layouts: q,k [nb*seq, heads*dp] (zero padded per head), vt [heads*dp, nb*seq]
    qp = torch.zeros(nb, seq, heads, dp, dtype=torch.float16, device=cuda)
    kp = torch.zeros_like(qp)
    qp[..., :d] = q.permute(0, 2, 1, 3)
    kp[..., :d] = k.permute(0, 2, 1, 3)
    vt = torch.zeros(heads, dp, nb, seq, dtype=torch.float16, device=cuda)
    vt[:, :d] = v.permute(1, 3, 0, 2)
    out = torch.full((nb * seq, heads * d), float("nan"), dtype=torch.float16, device=cuda)
    ops.attention(qp.reshape(nb * seq, heads * dp), kp.reshape(nb * seq, heads * dp),
                  vt.reshape(heads * dp, nb * seq), out, nb=nb, heads=heads, sq=seq, skv=seq, d_real=d, dp=dp,
                  k_bstride=seq, vt_bstride=seq)
    ref = F.scaled_dot_product_attention(q.float(), k.float(), v.float())  # (nb,heads,seq,d)
    ref = ref.permute(0, 2, 1, 3).reshape(nb * seq, heads * d)
    assert_close(out, ref, 2e-3, 4e-3, f"self-attn nb={nb} heads={heads} seq={seq} d={d}/{dp}")


@pytest.mark.parametrize("nb,heads,seq,d,dp,skv", [(1, 5, 4096, 64, 64, 77), (4, 20, 256, 64, 64, 77),
                                                    (2, 8, 1024, 40, 64, 77)])
def test_cross_attention_shared_kv(cuda, nb, heads, seq, d, dp, skv):
    """attn2: keys/values come from the 77-token prompt, identical for every batch item (prompt K/V cache)."""
    ops = _ops()
    q = _rand((nb, heads, seq, d), cuda, 1).half()
    k = _rand((heads, skv, d), cuda, 2).half()
    v = _rand((heads, skv, d), cuda, 3).half()
    qp = torch.zeros(nb, seq, heads, dp, dtype=torch.float16, device=cuda)
    qp[..., :d] = q.permute(0, 2, 1, 3)
    kp = torch.zeros(skv, heads, dp, dtype=torch.float16, device=cuda)
    kp[..., :d] = k.permute(1, 0, 2)
    vt_full = torch.zeros(heads, dp, 128, dtype=torch.float16, device=cuda)  # pitch 128, 77 valid columns
    vt_full[:, :d, :skv] = v.permute(0, 2, 1)
    vt = vt_full.reshape(heads * dp, 128)[:, :skv]
    out = torch.empty((nb * seq, heads * d), dtype=torch.float16, device=cuda)
    ops.attention(qp.reshape(nb * seq, heads * dp), kp.reshape(skv, heads * dp), vt, out, nb=nb, heads=heads, sq=seq,
                  skv=skv, d_real=d, dp=dp, k_bstride=0, vt_bstride=0)
    ref = F.scaled_dot_product_attention(q.float(), k.float()[None], v.float()[None])
    ref = ref.permute(0, 2, 1, 3).reshape(nb * seq, heads * d)
    assert_close(out, ref, 2e-3, 4e-3, f"cross-attn nb={nb} heads={heads} seq={seq}")


@pytest.mark.parametrize("nb,h,w,ca,cb,silu,eps", [
    (1, 64, 64, 320, 0, True, 1e-5), (2, 32, 32, 640, 0, False, 1e-6), (1, 16, 16, 1280, 1280, True, 1e-5),
    (1, 32, 32, 1280, 640, True, 1e-5),   # 1920 channels: groups of 60 straddle the concat boundary
    (4, 8, 8, 1280, 0, True, 1e-5), (1, 24, 24, 640, 320, True, 1e-5),
    (1, 64, 64, 640, 320, True, 1e-5),    # cluster of 8 CTAs per group
    (1, 8, 8, 1280, 1280, True, 1e-5), (1, 96, 96, 640, 320, True, 1e-5),   # too large for a cluster: whole-grid kernel
])
def test_groupnorm(cuda, nb, h, w, ca, cb, silu, eps):
    ops = _ops()
    xa = (_rand((nb, h, w, ca), cuda, 1) * 1.5 + 0.3).half()
    xb = (_rand((nb, h, w, cb), cuda, 2) * 0.7 - 0.2).half() if cb else None
    c = ca + cb
    gamma = (1 + 0.1 * _rand((c,), cuda, 3)).float()
    beta = (0.1 * _rand((c,), cuda, 4)).float()
    y = torch.empty((nb, h, w, c), dtype=torch.float16, device=cuda)
    ops.groupnorm(xa, xb, gamma, beta, y, eps=eps, silu=silu)
    x = xa if xb is None else torch.cat([xa, xb], dim=3)
    ref = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    assert_close(y, ref.permute(0, 2, 3, 1), 3e-3, 2e-3, f"groupnorm {ca}+{cb}")


@pytest.mark.parametrize("rows,c", [(4096, 320), (1024, 640), (77, 1280), (5, 64)])
def test_layernorm(cuda, rows, c):
    ops = _ops()
    x = (_rand((rows, c), cuda, 1) * 2 + 0.5).half()
    gamma = (1 + 0.1 * _rand((c,), cuda, 2)).float()
    beta = (0.1 * _rand((c,), cuda, 3)).float()
    y = torch.empty_like(x)
    ops.layernorm(x, gamma, beta, y)
    ref = F.layer_norm(x.float(), (c,), gamma, beta, 1e-5)
    assert_close(y, ref, 3e-3, 2e-3, f"layernorm {rows}x{c}")


def test_upsample2x(cuda):
    ops = _ops()
    x = _rand((2, 8, 12, 64), cuda, 1).half()
    y = torch.empty((2, 16, 24, 64), dtype=torch.float16, device=cuda)
    ops.upsample2x(x, y)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)


def test_smallconv_unet_conv_in(cuda):
    ops = _ops()
    x = _rand((2, 16, 16, 4), cuda, 1).half()
    w = _rand((320, 4, 3, 3), cuda, 2, 1 / 6.0).half()
    b = _rand((320,), cuda, 3).float()
    y = torch.empty((2, 16, 16, 320), dtype=torch.float16, device=cuda)
    ops.smallconv(x, w, b, y)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    assert_close(y, ref, 2e-3, 2e-3, "conv_in 4->320")


def test_smallconv_taesd_encoder_head_u8(cuda):
    """lib/pipeline.py:61-63 (u8 NHWC -> f32/255 -> NCHW) + VaeImageProcessor 2x-1 + EncoderTiny (x+1)/2 + conv."""
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    frame = torch.randint(0, 256, (1, 64, 48, 3), dtype=torch.uint8, generator=g).to(cuda)
    w = _rand((64, 3, 3, 3), cuda, 2, 1 / 5.0).half()
    b = _rand((64,), cuda, 3).float()
    y = torch.empty((1, 64, 48, 64), dtype=torch.float16, device=cuda)
    ops.smallconv(frame, w, b, y, flags=1)
    x = frame.float() / 255.0
    x = ((2 * x - 1) + 1) / 2
    ref = F.conv2d(x.permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    assert_close(y, ref, 2e-3, 2e-3, "taesd encoder head (u8 in)")


def test_smallconv_resize_nearest(cuda):
    """VaeImageProcessor.preprocess resizes (nearest) when the frame is not HxW (SURVEY a-4)."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    frame = torch.randint(0, 256, (1, 30, 40, 3), dtype=torch.uint8, generator=g).to(cuda)
    w = _rand((64, 3, 3, 3), cuda, 2, 1 / 5.0).half()
    y = torch.empty((1, 64, 64, 64), dtype=torch.float16, device=cuda)
    ops.smallconv(frame, w, None, y, flags=1)
    x = F.interpolate((frame.float() / 255.0).permute(0, 3, 1, 2), size=(64, 64))
    ref = F.conv2d(x, w.float(), None, padding=1).permute(0, 2, 3, 1)
    assert_close(y, ref, 2e-3, 2e-3, "encoder head with nearest resize")


def test_smallconv_taesd_decoder_head(cuda):
    ops = _ops()
    z = (_rand((1, 16, 16, 4), cuda, 1) * 2).half()
    w = _rand((64, 4, 3, 3), cuda, 2, 1 / 6.0).half()
    b = _rand((64,), cuda, 3).float()
    y = torch.empty((1, 16, 16, 64), dtype=torch.float16, device=cuda)
    ops.smallconv(z, w, b, y, flags=2 | 4)
    x = (torch.tanh(z.float() / 3) * 3).half().float()
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2), w.float(), b, padding=1)).permute(0, 2, 3, 1)
    assert_close(y, ref, 2e-3, 2e-3, "taesd decoder head")


@pytest.mark.parametrize("T", [1, 4])
def test_lcm_step_matches_streamdiffusion(cuda, T):
    """scheduler_step_batch + buffer update vs the oracle restatement (oracle/stream.py)."""
    ops = _ops()
    from oracle import stream as ostream
    hw = (8, 8)
    t_list = [18, 26, 35, 45][:T] if T > 1 else [32]
    so = ostream.StreamOracle({}, None, {}, t_list, 64, 64)
    so.prepare(torch.zeros(1, 77, 8), guidance_scale=0.0)
    x = _rand((T, 4, *hw), "cpu", 1).half().float()
    eps = _rand((T, 4, *hw), "cpu", 2).half().float()
    so.init_noise = so.init_noise.half().float()
    x0 = so.scheduler_step_batch(eps, x)
    coef = torch.stack([so.alpha_prod_t_sqrt.flatten(), so.beta_prod_t_sqrt.flatten(), so.c_skip.flatten(),
                        so.c_out.flatten()]).float().contiguous().to(cuda)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().half().to(cuda)
    xd, ed, nd = nhwc(x), nhwc(eps), nhwc(so.init_noise)
    outd = torch.empty((1, *hw, 4), dtype=torch.float16, device=cuda)
    ops.lcm_step(xd, ed, nd, coef, outd)
    assert_close(outd[0].permute(2, 0, 1), x0[-1], 2e-3, 2e-3, "x0 of the last slot")
    if T > 1:
        buf = so.alpha_prod_t_sqrt[1:] * x0[:-1] + so.beta_prod_t_sqrt[1:] * so.init_noise[1:]
        assert_close(xd[1:].permute(0, 3, 1, 2), buf, 2e-3, 2e-3, "x_t_latent_buffer")


def test_post_u8_truncation_semantics(cuda):
    """fp16 chain of DecoderTiny tail, postprocess_image and lib/pipeline.py:72-74: bit exact vs torch half ops."""
    ops = _ops()
    y = (_rand((1, 32, 32, 3), cuda, 1) * 0.4 + 0.5).half()
    y[0, 0, 0, 0] = 1.7   # clamps
    y[0, 0, 1, 0] = -0.3
    out = torch.empty((1, 3, 32, 32), dtype=torch.uint8, device=cuda)
    ops.post_u8(y, out)
    img = y.permute(0, 3, 1, 2).mul(2).sub(1)           # fp16
    den = (img / 2 + 0.5).clamp(0, 1)                   # fp16
    ref = (den * 255.0).clamp(0, 255).to(torch.uint8)   # truncation
    assert torch.equal(out, ref), f"mismatch {(out != ref).sum().item()} px"


# ---- codec boundary: NV12 <-> RGB colour conversion next to NVDEC / NVENC -----------------------------------
def _csc_ref(flags):
    kr, kb = (0.299, 0.114) if flags & 1 else (0.2126, 0.0722)
    yo, ys, cs = (0.0, 1.0, 1.0) if flags & 2 else (16.0, 219.0 / 255.0, 224.0 / 255.0)
    return kr, 1.0 - kr - kb, kb, yo, ys, cs


@pytest.mark.parametrize("flags", [0, 1, 2, 3])
@pytest.mark.parametrize("h,w", [(64, 96), (512, 512), (30, 50)])
def test_nv12_rgb_colour_conversion(cuda, flags, h, w):
    """Both directions against the textbook BT.709 / BT.601 matrices (limited and full range), 2x2 chroma averaging."""
    from ai_rtc_agent_b200.host import codec
    g = torch.Generator().manual_seed(7)
    rgb = torch.randint(0, 256, (1, 3, h, w), dtype=torch.uint8, generator=g)
    kr, kg, kb, yo, ys, cs = _csc_ref(flags)
    r, gg, b = (rgb[0, i].double() for i in range(3))
    yl = kr * r + kg * gg + kb * b
    y_ref = (yo + ys * yl).round().clamp(0, 255)
    cb = (b - yl) / (2 * (1 - kb))
    cr = (r - yl) / (2 * (1 - kr))
    pool = lambda t: torch.nn.functional.avg_pool2d(t[None, None], 2, ceil_mode=True, count_include_pad=False)[0, 0]
    cb_ref = (128 + cs * pool(cb)).round().clamp(0, 255)
    cr_ref = (128 + cs * pool(cr)).round().clamp(0, 255)
    y, uv = codec.rgb_to_nv12(rgb.to(cuda), flags)
    assert (y.cpu().double() - y_ref).abs().max() <= 1
    assert (uv.cpu()[:, 0::2].double()[:, :cb_ref.shape[1]] - cb_ref).abs().max() <= 1
    assert (uv.cpu()[:, 1::2].double()[:, :cr_ref.shape[1]] - cr_ref).abs().max() <= 1
    # decode direction: exact formula on the encoder's planes
    back = codec.nv12_to_rgb(y, uv, flags).cpu()[0].double()      # (H,W,3)
    yy = (y.cpu().double() - yo) / ys
    up = lambda t: t.repeat_interleave(2, 0).repeat_interleave(2, 1)[:h, :w]
    cbd = up((uv.cpu()[:, 0::2].double() - 128) / cs)
    crd = up((uv.cpu()[:, 1::2].double() - 128) / cs)
    rr = yy + 2 * (1 - kr) * crd
    bb = yy + 2 * (1 - kb) * cbd
    gr = (yy - kr * rr - kb * bb) / kg
    ref = torch.stack([rr, gr, bb], dim=-1).round().clamp(0, 255)
    assert (back - ref).abs().max() <= 1
    # a smooth image survives the round trip closely (chroma sub-sampling aside)
    xs = torch.linspace(0, 255, w)[None, :].expand(h, w)
    smooth = torch.stack([xs, xs.flip(1), torch.full_like(xs, 128.0)]).round().to(torch.uint8)[None]
    ys_, uvs = codec.rgb_to_nv12(smooth.to(cuda), flags)
    rt = codec.nv12_to_rgb(ys_, uvs, flags).cpu()[0].permute(2, 0, 1).double()
    assert (rt - smooth[0].double()).abs().max() <= 6     # 2x2 chroma averaging of a 5-levels-per-pixel ramp + two roundings


def test_codec_sessions_report_unavailable(cuda):
    from ai_rtc_agent_b200.host import codec
    libs = codec.codec_libraries()
    assert set(libs) == {"nvdec", "nvenc"}
    with pytest.raises(codec.CodecUnavailable):
        codec.open_decoder()
    with pytest.raises(codec.CodecUnavailable):
        codec.open_encoder()

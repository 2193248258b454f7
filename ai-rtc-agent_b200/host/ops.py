"""Thin torch-tensor wrappers over the operator-level C ABI (used by tests and tools; the per-frame
engine calls the same kernels from C++ without going through Python)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import capi


_SCRATCH = {}  # split-K workspaces kept alive between calls (stream-ordered reuse)


def pack_conv_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """[O,I,kh,kw] -> [O, kh*kw*I] fp16 with K order [tap][c] (tap = ky*kw + kx)."""
    o, i, kh, kw = w_oihw.shape
    return w_oihw.permute(0, 2, 3, 1).reshape(o, kh * kw * i).contiguous().to(torch.float16)


def _view(t: torch.Tensor) -> capi.ActView:
    assert t.dtype == torch.float16 and t.is_cuda and t.dim() == 4
    n, h, w, c = t.shape
    assert t.stride(3) == 1 and t.stride(2) % 8 == 0
    ld = t.stride(2)
    assert t.stride(1) == w * ld and t.stride(0) == h * w * ld, "NHWC view must be dense in n,h,w"
    return capi.ActView(t.data_ptr(), n, h, w, c, ld)


def igemm(srcs: Sequence[Tuple[torch.Tensor, int]], w: torch.Tensor, out: torch.Tensor, *,
          stride: int = 1, colbias: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
          acc_scale: float = 1.0, res_scale: float = 1.0, relu: bool = False, geglu: bool = False,
          bn: int = 0, splits: int = 1, n_valid: Optional[int] = None, timeline: Optional[torch.Tensor] = None, swap: bool = False,
          tconv: bool = False, pair: bool = False,
          rowstat_out: Optional[torch.Tensor] = None, rowstat_in: Optional[torch.Tensor] = None,
          colsum: Optional[torch.Tensor] = None, ln_c: int = 0, ln_eps: float = 1e-5,
          out2: Optional[torch.Tensor] = None, col2: int = 0) -> torch.Tensor:
    """srcs: [(NHWC fp16 tensor, ntap)], w: packed fp16 [rows, K]; out: NHWC fp16 [nb,ho,wo,ldc>=n]."""
    d = capi.IgemmDesc()
    d.nseg = len(srcs)
    for i, (t, ntap) in enumerate(srcs):
        d.src[i] = _view(t)
        d.ntap[i] = ntap
    assert w.dtype == torch.float16 and w.is_contiguous()
    d.w, d.w_rows, d.w_ld = w.data_ptr(), w.shape[0], w.shape[1]
    d.stride = stride
    nb, ho, wo, _ = out.shape
    d.nb, d.ho, d.wo = nb, ho, wo
    d.bn, d.splits = bn, splits
    d.swap = int(swap)
    d.out, d.ldc = out.data_ptr(), out.stride(2)
    nv = n_valid if n_valid is not None else out.shape[3]
    d.n_valid = nv
    if timeline is not None:
        d.partial = timeline.data_ptr()   # debug: int64 tensor [ctas, 8] receiving globaltimer stamps (tap kernel only)
    if colbias is not None:
        assert colbias.dtype == torch.float32 and colbias.is_contiguous()
        d.colbias = colbias.data_ptr()
        d.colbias_bstride = colbias.shape[1] if colbias.dim() == 2 and colbias.shape[0] > 1 else 0
    if res is not None:
        assert res.dtype == torch.float16
        d.res, d.ldr = res.data_ptr(), res.stride(2)
    d.acc_scale, d.res_scale = acc_scale, res_scale
    d.flags = (capi.IG_RELU if relu else 0) | (capi.IG_GEGLU if geglu else 0) | (capi.IG_TCONV if tconv else 0) | (capi.IG_PAIR if pair else 0)
    if rowstat_out is not None:
        assert rowstat_out.dtype == torch.int64 and rowstat_out.is_contiguous()
        d.rowstat_out = rowstat_out.data_ptr()
    if rowstat_in is not None:
        assert rowstat_in.dtype == torch.int64 and colsum is not None and colsum.dtype == torch.float32 and ln_c > 0
        d.rowstat_in, d.colsum, d.ln_c, d.ln_eps = rowstat_in.data_ptr(), colsum.data_ptr(), ln_c, ln_eps
    if out2 is not None:
        assert out2.dtype == torch.float16 and out2.dim() == 2 and out2.stride(1) == 1
        d.out2, d.ld2, d.col2 = out2.data_ptr(), out2.stride(0), col2
    capi.check(capi.lib().b2sd_op_igemm(C.byref(d), capi.current_stream_ptr()), "b2sd_op_igemm")
    return out


def _sp():
    return capi.current_stream_ptr()


def attention(q, k, vt, out, *, nb, heads, sq, skv, d_real, dp, k_bstride, vt_bstride):
    """q [nb*sq, >=heads*dp], k [rows, >=heads*dp], vt [heads*dp, cols] (2-D fp16 views), out [nb*sq, heads*d_real]."""
    d = capi.AttnDesc()
    d.q, d.ldq = q.data_ptr(), q.stride(0)
    d.k, d.ldk, d.k_bstride, d.k_rows = k.data_ptr(), k.stride(0), k_bstride, k.shape[0]
    d.vt, d.ldvt, d.vt_bstride, d.vt_cols = vt.data_ptr(), vt.stride(0), vt_bstride, vt.shape[1]
    d.out, d.ldo = out.data_ptr(), out.stride(0)
    d.nb, d.heads, d.sq, d.skv, d.d_real, d.dp = nb, heads, sq, skv, d_real, dp
    capi.check(capi.lib().b2sd_op_attention(C.byref(d), _sp()), "b2sd_op_attention")
    return out


def groupnorm(xa, xb, gamma, beta, y, *, groups=32, eps=1e-5, silu=True):
    """xa/xb: NHWC fp16 (xb may be None); y: NHWC fp16 with C = ca + cb."""
    nb, h, w, ca = xa.shape
    cb = 0 if xb is None else xb.shape[3]
    capi.check(capi.lib().b2sd_op_groupnorm(
        xa.data_ptr(), ca, xa.stride(2), 0 if xb is None else xb.data_ptr(), cb, 0 if xb is None else xb.stride(2),
        gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), y.stride(2), nb, h * w, groups, eps, int(silu), _sp()),
        "b2sd_op_groupnorm")
    return y


def layernorm(x, gamma, beta, y, eps=1e-5):
    rows, c = x.shape
    capi.check(capi.lib().b2sd_op_layernorm(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(),
                                            y.data_ptr(), y.stride(0), rows, c, eps, _sp()), "b2sd_op_layernorm")
    return y


def upsample2x(x, y):
    nb, h, w, c = x.shape
    capi.check(capi.lib().b2sd_op_upsample2x(x.data_ptr(), y.data_ptr(), nb, h, w, c, _sp()), "b2sd_op_upsample2x")
    return y


def smallconv(x, w_oihw, bias, y, *, flags=0):
    """x: NHWC fp16 (or u8 when flags&1) [nb,in_h,in_w,cin]; y NHWC fp16 [nb,h,w,cout]."""
    nb, in_h, in_w, cin = x.shape
    _, h, w, cout = y.shape
    capi.check(capi.lib().b2sd_op_smallconv(x.data_ptr(), w_oihw.data_ptr(), 0 if bias is None else bias.data_ptr(),
                                            y.data_ptr(), y.stride(2), nb, h, w, cin, cout, in_h, in_w, flags, _sp()),
               "b2sd_op_smallconv")
    return y


def lcm_step(x, eps, noise, coef, out_latent, do_add_noise=True):
    T, hw = x.shape[0], x.shape[1] * x.shape[2]
    capi.check(capi.lib().b2sd_op_lcm_step(x.data_ptr(), eps.data_ptr(), noise.data_ptr(), coef.data_ptr(),
                                           out_latent.data_ptr(), T, hw, int(do_add_noise), _sp()), "b2sd_op_lcm_step")
    return out_latent


def post_u8(y, out):
    nb, h, w, _ = y.shape
    capi.check(capi.lib().b2sd_op_post_u8(y.data_ptr(), y.stride(2), out.data_ptr(), nb, h, w, _sp()), "b2sd_op_post_u8")
    return out

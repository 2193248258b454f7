"""Thin torch-tensor wrappers over the operator-level C ABI (used by tests and tools; the per-frame
engine calls the same kernels from C++ without going through Python)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import capi


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
          bn: int = 0, splits: int = 1, n_valid: Optional[int] = None) -> torch.Tensor:
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
    d.out, d.ldc = out.data_ptr(), out.stride(2)
    nv = n_valid if n_valid is not None else out.shape[3]
    d.n_valid = nv
    keep = []
    if splits > 1:
        nfl = capi.lib().b2sd_igemm_partial_floats(splits, nb * ho * wo, nv)
        part = torch.empty(nfl, dtype=torch.float32, device=out.device)
        keep.append(part)
        d.partial = part.data_ptr()
    if colbias is not None:
        assert colbias.dtype == torch.float32 and colbias.is_contiguous()
        d.colbias = colbias.data_ptr()
        d.colbias_bstride = colbias.shape[1] if colbias.dim() == 2 and colbias.shape[0] > 1 else 0
    if res is not None:
        assert res.dtype == torch.float16
        d.res, d.ldr = res.data_ptr(), res.stride(2)
    d.acc_scale, d.res_scale = acc_scale, res_scale
    d.flags = (capi.IG_RELU if relu else 0) | (capi.IG_GEGLU if geglu else 0)
    capi.check(capi.lib().b2sd_op_igemm(C.byref(d), capi.current_stream_ptr()), "b2sd_op_igemm")
    if keep:
        torch.cuda.current_stream().synchronize()
    return out

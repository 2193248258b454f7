"""Shared helpers for the parity tests."""
import torch


def describe_mismatch(got: torch.Tensor, ref: torch.Tensor, tol_abs: float, tol_rel: float) -> str:
    """Compact remote-debuggable summary: error stats + where (row%8 / col%64 patterns) it is wrong."""
    got = got.float().cpu()
    ref = ref.float().cpu()
    err = (got - ref).abs()
    bad = err > (tol_abs + tol_rel * ref.abs())
    lines = [f"shape={tuple(got.shape)} max_err={err.max().item():.4g} ref_absmax={ref.abs().max().item():.4g} "
             f"bad={bad.sum().item()}/{bad.numel()} nan_got={torch.isnan(got).sum().item()}"]
    if bad.any():
        idx = bad.nonzero()
        lines.append("first bad idx: " + str(idx[:6].tolist()))
        for k in range(min(6, idx.shape[0])):
            t = tuple(idx[k].tolist())
            lines.append(f"  {t}: got={got[t].item():.5g} ref={ref[t].item():.5g}")
        flat = bad.reshape(-1, bad.shape[-1])
        rows_bad = flat.any(dim=1)
        cols_bad = flat.any(dim=0)
        lines.append(f"rows bad {rows_bad.sum().item()}/{rows_bad.numel()} cols bad {cols_bad.sum().item()}/{cols_bad.numel()}")
        rb = rows_bad.nonzero().flatten()
        cb = cols_bad.nonzero().flatten()
        lines.append("bad rows (first 24): " + str(rb[:24].tolist()))
        lines.append("bad cols (first 24): " + str(cb[:24].tolist()))
    return "\n".join(lines)


def assert_close(got, ref, tol_abs, tol_rel, what=""):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    err = (got - ref).abs()
    ok = bool((err <= tol_abs + tol_rel * ref.abs()).all()) and not bool(torch.isnan(got).any())
    assert ok, what + "\n" + describe_mismatch(got, ref, tol_abs, tol_rel)

"""Event-timed micro-benchmarks of individual launches (L2-warm, 50 iterations each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_rtc_agent_b200.host import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device=dev) * scale).half()
def timeit(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1000
def conv_case(nb, h, w, cin, cout, bn, splits, res=False, relu=False):
    x = rnd(nb, h, w, cin); wt = ops.pack_conv_weight(rnd(cout, cin, 3, 3, scale=(9*cin) ** -0.5)); b = torch.randn(1, cout, device=dev)
    y = torch.empty(nb, h, w, cout, device=dev, dtype=torch.float16)
    if splits > 1:
        nfl = ops.capi.lib().b2sd_igemm_partial_floats(splits, nb*h*w, cout)
    us = timeit(lambda: ops.igemm([(x, 9)], wt, y, colbias=b, bn=bn, splits=splits, res=x if res else None, relu=relu))
    fl = 2.0 * nb * h * w * cout * cin * 9
    print(f"conv {nb}x{h}x{w} {cin}->{cout} bn={bn} splits={splits}: {us:8.1f} us  {fl/us/1e6:7.1f} TF/s")
import itertools
if "sweep" in sys.argv:
    for (h, c) in [(64, 320), (32, 640), (16, 1280), (8, 1280)]:
        for bn, sp in itertools.product([64, 128, 160, 256], [1, 2, 4, 8]):
            if c % bn: continue
            try: conv_case(1, h, h, c, c, bn, sp)
            except Exception as e: print("fail", h, c, bn, sp, str(e)[:80])
    sys.exit(0)
for args in [(1,512,512,64,64,64,1,True,True), (1,256,256,64,64,64,1,True,True), (1,64,64,320,320,64,1), (1,64,64,320,320,160,1),
             (1,32,32,640,640,64,1), (1,32,32,640,640,64,2), (1,32,32,640,640,128,2), (1,16,16,1280,1280,64,1), (1,16,16,1280,1280,64,4),
             (1,16,16,1280,1280,128,8), (1,16,16,1280,1280,256,8), (1,8,8,1280,1280,64,8), (1,8,8,1280,1280,64,16), (1,16,16,2560,1280,64,4)]:
    conv_case(*args)

def gn_case(nb, h, w, c, cb=0):
    xa = rnd(nb, h, w, c); xb = rnd(nb, h, w, cb) if cb else None
    g = torch.ones(c + cb, device=dev); b = torch.zeros(c + cb, device=dev); y = torch.empty(nb, h, w, c + cb, device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.groupnorm(xa, xb, g, b, y))
    print(f"groupnorm {nb}x{h}x{w}x{c}+{cb}: {us:7.1f} us  ({(xa.numel() + (xb.numel() if cb else 0)) * 4 / us / 1e3:6.1f} GB/s r+w)")
def ln_case(rows, c):
    x = rnd(rows, c); g = torch.ones(c, device=dev); b = torch.zeros(c, device=dev); y = torch.empty_like(x)
    us = timeit(lambda: ops.layernorm(x, g, b, y))
    print(f"layernorm {rows}x{c}: {us:7.1f} us")
def attn_case(nb, heads, seq, skv=None):
    skv = skv or seq
    qk = rnd(nb * seq, 2 * heads * 64); vt = rnd(heads * 64, nb * max(skv, 128)); o = torch.empty(nb * seq, heads * 64, device=dev, dtype=torch.float16)
    k = qk[:, heads * 64:]
    us = timeit(lambda: ops.attention(qk[:, :heads * 64], k, vt[:, :nb * skv] if skv == seq else vt[:, :skv], o, nb=nb, heads=heads, sq=seq, skv=skv,
                                      d_real=64, dp=64, k_bstride=seq if skv == seq else 0, vt_bstride=seq if skv == seq else 0))
    fl = 4.0 * nb * heads * seq * skv * 64
    print(f"attention nb={nb} heads={heads} seq={seq} skv={skv}: {us:7.1f} us {fl/us/1e6:7.1f} TF/s")
def lin_case(m, k, n, bn=0, geglu=False):
    x = rnd(1, 1, m, k); w = rnd(n * (2 if geglu else 1), k, scale=k ** -0.5); b = torch.randn(1, n * (2 if geglu else 1), device=dev)
    y = torch.empty(1, 1, m, n, device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.igemm([(x, 1)], w, y, colbias=b, bn=bn, geglu=geglu, n_valid=n))
    fl = 2.0 * m * k * n * (2 if geglu else 1)
    print(f"linear m={m} k={k} n={n} bn={bn} geglu={geglu}: {us:7.1f} us {fl/us/1e6:7.1f} TF/s")
def head_case():
    fr = torch.randint(0, 256, (1, 512, 512, 3), dtype=torch.uint8, device=dev); w = rnd(64, 3, 3, 3); b = torch.randn(64, device=dev)
    y = torch.empty(1, 512, 512, 64, device=dev, dtype=torch.float16)
    print(f"smallconv head 512x512 3->64: {timeit(lambda: ops.smallconv(fr, w, b, y, flags=1)):7.1f} us")
    x = rnd(1, 64, 64, 4); w2 = rnd(320, 4, 3, 3); b2 = torch.randn(320, device=dev); y2 = torch.empty(1, 64, 64, 320, device=dev, dtype=torch.float16)
    print(f"smallconv conv_in 64x64 4->320: {timeit(lambda: ops.smallconv(x, w2, b2, y2)):7.1f} us")
if "all" in sys.argv or len(sys.argv) == 1:
    gn_case(1, 64, 64, 320); gn_case(1, 64, 64, 640, 320); gn_case(1, 32, 32, 640); gn_case(1, 16, 16, 1280); gn_case(1, 8, 8, 1280, 1280)
    ln_case(4096, 320); ln_case(1024, 640); ln_case(256, 1280); ln_case(64, 1280)
    attn_case(1, 5, 4096); attn_case(1, 10, 1024); attn_case(1, 20, 256); attn_case(1, 20, 64); attn_case(1, 5, 4096, 77); attn_case(1, 20, 256, 77)
    lin_case(4096, 320, 320); lin_case(4096, 320, 640); lin_case(4096, 320, 1280, 128, True); lin_case(4096, 1280, 320); lin_case(1024, 640, 640); lin_case(1024, 640, 2560, 128, True)
    lin_case(1024, 2560, 640); lin_case(256, 1280, 1280); lin_case(256, 1280, 5120, 128, True); lin_case(256, 5120, 1280); lin_case(64, 1280, 1280)
    head_case()

"""Per-CTA timeline of the tap-by-tap igemm kernel (globaltimer stamps): where does a CTA's time go?"""
import os, sys
pass
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_rtc_agent_b200.host import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device=dev) * scale).half()
def run(nb, h, w, cin, cout, bn, splits, res):
    x = rnd(nb, h, w, cin); wt = ops.pack_conv_weight(rnd(cout, cin, 3, 3, scale=(9*cin) ** -0.5)); b = torch.randn(1, cout, device=dev)
    y = torch.empty(nb, h, w, cout, device=dev, dtype=torch.float16)
    for _ in range(3): ops.igemm([(x, 9)], wt, y, colbias=b, bn=bn, splits=splits, res=x if res else None)
    ts = torch.zeros(65536, 8, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ops.igemm([(x, 9)], wt, y, colbias=b, bn=bn, splits=splits, res=x if res else None, timeline=ts)
    torch.cuda.synchronize()
    t = ts.cpu()
    t = t[t[:, 0] > 0].double()
    t0 = t[:, 0].min()
    names = ["entry", "prologue done", "first TMA issued", "first stage landed", "last MMA issued", "accum ready", "epilogue done", "exit"]
    print(f"--- conv {nb}x{h}x{w} {cin}->{cout} bn={bn} splits={splits}: {t.shape[0]} CTAs, kernel span {(t[:, 7].max() - t0) / 1000:.1f} us")
    for i in range(1, 8):
        d = (t[:, i] - t[:, i - 1]) / 1000
        print(f"  {names[i-1]:20s} -> {names[i]:20s}: median {d.median():6.2f} us  p90 {d.quantile(0.9):6.2f}  max {d.max():6.2f}")
    d = (t[:, 7] - t[:, 0]) / 1000
    print(f"  CTA lifetime median {d.median():.2f} us; CTA starts: first {0:.1f} .. last {(t[:, 0].max() - t0) / 1000:.1f} us")
splits = int(os.getenv("SPL", "1"))
run(1, 512, 512, 64, 64, 64, 1, True)
run(1, 64, 64, 320, 320, 64, splits, False)
run(1, 16, 16, 1280, 1280, 64, 4, False)

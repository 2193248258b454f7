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
for args in [(1,512,512,64,64,64,1,True,True), (1,256,256,64,64,64,1,True,True), (1,64,64,320,320,64,1), (1,64,64,320,320,160,1),
             (1,32,32,640,640,64,1), (1,32,32,640,640,64,2), (1,32,32,640,640,128,2), (1,16,16,1280,1280,64,1), (1,16,16,1280,1280,64,4),
             (1,16,16,1280,1280,128,8), (1,16,16,1280,1280,256,8), (1,8,8,1280,1280,64,8), (1,8,8,1280,1280,64,16), (1,16,16,2560,1280,64,4)]:
    conv_case(*args)

"""Where does the fixed per-launch cost go?  A chain of dependent igemm launches replayed from one CUDA graph, every CTA
stamping globaltimer at 8 points (needs the instrumented build: make -C ai-rtc-agent_b200/csrc timeline;
B200SD_LIB=ai-rtc-agent_b200/libb200sd_tl.so python tools/timeline_chain.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_rtc_agent_b200.host import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def rnd(*s, scale=1.0): return (torch.randn(*s, device=dev) * scale).half()
NAMES = ["entry", "prologue+pdl_wait", "first TMA issued", "first stage landed", "last MMA issued", "tile-0 epilogue", "epilogue end", "exit"]
def chain(h, cin, taps, bn, splits, swap, n=8):
    c = cin
    xs = [rnd(1, h, h, c) for _ in range(2)]
    if taps == 9: ws = [ops.pack_conv_weight(rnd(c, c, 3, 3, scale=(9 * c) ** -0.5)) for _ in range(n)]
    else: ws = [rnd(c, c, scale=c ** -0.5) for _ in range(n)]
    b = torch.zeros(1, c, device=dev)
    ts = torch.zeros(n, 4096, 8, dtype=torch.int64, device=dev)
    def run(tl):
        for i in range(n):   # ping-pong: launch i reads what launch i-1 wrote
            ops.igemm([(xs[i & 1], taps)], ws[i], xs[(i + 1) & 1], colbias=b, bn=bn, splits=splits, swap=swap, timeline=ts[i] if tl else None)
    run(False); torch.cuda.synchronize()
    st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            run(True)
    g.replay(); torch.cuda.synchronize(); ts.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    t = ts.cpu().double()
    print(f"=== {'swap' if swap else 'base'} {h}x{h} {c}->{c} taps={taps} bn={bn} splits={splits}: chain of {n}: {e0.elapsed_time(e1) * 1000 / n:.2f} us per launch")
    prev_exit = None
    for i in range(n):
        ti = t[i]; ti = ti[ti[:, 0] > 0]
        ent, ex = ti[:, 0].min(), ti[:, 7].max()
        gap = (ent - prev_exit) / 1000 if prev_exit is not None else float("nan")
        hand = (ti[:, 1].min() - prev_exit) / 1000 if prev_exit is not None else float("nan")   # last exit of i-1 -> first pdl_wait return of i
        crit = (ex - ti[:, 1].min()) / 1000                                                     # first pdl_wait return -> last exit
        ph = [((ti[:, k] - ti[:, k - 1]).median() / 1000).item() for k in range(1, 8)]
        print(f" launch {i}: {ti.shape[0]:4d} CTAs  gap(prev last exit -> first entry) {gap:6.2f} us | handoff {hand:5.2f} | wait->last exit {crit:6.2f} us | "
              + " ".join(f"{p:5.2f}" for p in ph))
        prev_exit = ex
    print("   phases: " + " | ".join(f"{NAMES[k-1]}->{NAMES[k]}" for k in range(1, 8)))
print("split launches: last three phases are  staged -> cluster barrier passed -> reduced (+stored)")
if "split" not in sys.argv:
    chain(64, 320, 1, 64, 1, False)
    chain(64, 320, 9, 64, 1, False)
chain(16, 1280, 1, 64, 4, False)
chain(16, 1280, 9, 64, 4, False)
chain(16, 1280, 9, 256, 8, True)
chain(8, 1280, 9, 64, 8, True)

"""Stream-K attention (B2_ATTN_SK=1) against the plain kernel: max |diff| and event-timed duration (GPU box)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_rtc_agent_b200.host import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
heads, seq = 5, 4096
qk = torch.randn(seq, 2 * heads * 64, device=dev).half(); vt = torch.randn(heads * 64, seq, device=dev).half()
o = torch.empty(seq, heads * 64, device=dev, dtype=torch.float16)
def run():
    ops.attention(qk[:, :heads * 64], qk[:, heads * 64:], vt, o, nb=1, heads=heads, sq=seq, skv=seq, d_real=64, dp=64, k_bstride=seq, vt_bstride=seq)
run(); torch.cuda.synchronize()
q = qk[:, :heads * 64].float().view(seq, heads, 64).transpose(0, 1); k = qk[:, heads * 64:].float().view(seq, heads, 64).transpose(0, 1)
v = vt.float().view(heads, 64, seq).transpose(1, 2)
ref = torch.softmax(q @ k.transpose(1, 2) / 8.0, dim=-1) @ v
err = (o.float().view(seq, heads, 64).transpose(0, 1) - ref).abs().max().item()
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): run()
e1.record(); torch.cuda.synchronize()
print(f"B2_ATTN_SK={os.environ.get('B2_ATTN_SK')}: max|diff| vs fp32 reference {err:.3e}, {e0.elapsed_time(e1) / 30 * 1000:.1f} us per launch")

"""one attention launch per head count (for ncu): python tools/attn_one.py 1 5"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_rtc_agent_b200.host import ops
dev = torch.device("cuda:0"); torch.manual_seed(0)
for heads in [int(a) for a in sys.argv[1:]] or [5]:
    seq = 4096
    qk = (torch.randn(seq, 2 * heads * 64, device=dev)).half(); vt = torch.randn(heads * 64, seq, device=dev).half()
    o = torch.empty(seq, heads * 64, device=dev, dtype=torch.float16)
    for _ in range(2):
        ops.attention(qk[:, :heads * 64], qk[:, heads * 64:], vt, o, nb=1, heads=heads, sq=seq, skv=seq, d_real=64, dp=64, k_bstride=seq, vt_bstride=seq)
    torch.cuda.synchronize()

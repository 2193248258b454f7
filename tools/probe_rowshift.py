import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_rtc_agent_b200.host import capi
lib = capi.lib()
lib.b2sd_probe_umma_rowshift.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
torch.manual_seed(0)
rows = 256
A = torch.randn(rows, 64, device=dev).half(); B = torch.randn(64, 64, device=dev).half()
for pitch in (8, 10, 18):
    for mode in (0, 1):
        res = []
        for shift in range(0, 12):
            D = torch.zeros(128, 64, device=dev)
            if (15 * pitch + 7 + shift) >= rows: res.append("--"); continue
            rc = lib.b2sd_probe_umma_rowshift(A.data_ptr(), rows, B.data_ptr(), D.data_ptr(), shift, pitch, mode, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            idx = torch.tensor([(r // 8) * pitch + r % 8 + shift for r in range(128)], device=dev)
            ref = A[idx].float() @ B.float().t()
            err = (D - ref).abs().max().item()
            res.append("ok" if err < 1e-2 else f"{err:.1f}")
        print(f"pitch={pitch:2d} base_mode={mode}: " + " ".join(res))

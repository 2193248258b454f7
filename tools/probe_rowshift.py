import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
PROBE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe", "libb200probe.so")  # make -C tools/probe
lib = C.CDLL(PROBE)
lib.b2sd_probe_umma_rowshift.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
dev = torch.device("cuda:0")
torch.manual_seed(0)
rows = 256
A = torch.randn(rows, 64, device=dev).half(); B = torch.randn(64, 64, device=dev).half()
tm = torch.zeros(2, dtype=torch.int64, device=dev)
for mode, what in [(128, "uniform issue, plain"), (176, "uniform, wait+fence per 4 MMAs"), (256+128, "unrolled x4, plain"), (256+128+48, "unrolled x4, wait+fence"), (256+128+48+4, "unrolled x4, wait+fence+commit")]:
    D = torch.zeros(128, 64, device=dev)
    lib.b2sd_probe_umma_rowshift(A.data_ptr(), rows, B.data_ptr(), D.data_ptr(), 0, 8, mode, torch.cuda.current_stream().cuda_stream, tm.data_ptr())
    torch.cuda.synchronize()
    print(f"{what:40s}: {tm[0].item() / 1024:7.1f} cycles, {tm[1].item() / 1024:7.1f} ns per 128x64x16 MMA")

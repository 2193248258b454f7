import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200SD_DEBUG_SYNC", "1")
os.environ.setdefault("B200SD_TEST_GRAPH", "0")
from tests.test_engine_gpu import _build
from oracle import weights as ow
turbo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tl = [18, 26, 35, 45] if len(sys.argv) <= 2 else [int(v) for v in sys.argv[2].split(",")]
sd, orc = _build("tiny", bool(turbo), tl, 128, torch.device("cuda:0"))
print("launches", sd.launches_per_step)
try:
    out = sd.step_u8(ow.make_frame(128, 128, seed=0).cuda())
    torch.cuda.synchronize()
    print("ok", out.float().mean().item())
except Exception as e:
    print("FAILED:", e)

"""Per-launch device-time table of one SD-Turbo 512x512 frame (eager replay with CUDA events).
Usage (GPU box): python tools/profile_ops.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("B200SD_SYNTHETIC_WEIGHTS", "1")
os.environ["NVENC"] = "1"
import torch
from ai_rtc_agent_b200.host.pipeline import StreamDiffusionPipeline
model = os.getenv("B200SD_MODEL", "stabilityai/sd-turbo")
tl = [int(v) for v in os.getenv("B200SD_T", "32").split(",")]
hw = int(os.getenv("B200SD_HW", "512"))
pipe = StreamDiffusionPipeline(model, t_index_list=tl, width=hw, height=hw)
frame = torch.randint(0, 256, (1, hw, hw, 3), dtype=torch.uint8).cuda()
for _ in range(3):
    pipe(frame)
prof = pipe.model.stream.profile(frame, iters=5)
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ops_profile.json"
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(prof, open(out, "w"), indent=0)
tot = sum(o["ms"] for o in prof)
print(f"total eager {tot:.3f} ms over {len(prof)} launches")
for o in sorted(prof, key=lambda o: -o["ms"])[:60]:
    tf = o["flops"] / (o["ms"] * 1e-3) / 1e12 if o["ms"] > 0 else 0
    print(f"{o['ms']*1000:8.1f} us {tf:7.1f} TF/s  {o['name']}")

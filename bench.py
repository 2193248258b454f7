#!/usr/bin/env python
"""Headline benchmark (BASELINE.json metric): frames/sec at 512x512 SD-Turbo 1-step img2img, stream-batch 1,
synthetic RGB frame feed, one independent video stream per GPU (weak scaling, NCCL weight broadcast at init
only, no per-step collective).

  python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a engine through the public API)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (fp32 oracle on host cores)

A "step" is one frame through StreamDiffusionPipeline.__call__.  `value` is timed on the device with the
input frame already resident in HBM; `e2e` includes, every step, the pinned-host -> device copy of the frame
and the device -> pinned-host read of the result.  Prints ONE JSON line on rank 0."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload table: BASELINE.json configs[1] is the default (the configuration the headline metric is quoted on);
# the others are the remaining single-GPU configs, selectable for additional measurements (--workload)
WORKLOADS = {
    "sd-turbo-512": dict(model="stabilityai/sd-turbo", t=[32], hw=512, gflop=1068.0,
                         metric="frames/sec at 512x512 SD-Turbo img2img (1-step, stream-batch 1)",
                         name="SD-Turbo 1-step img2img 512x512, stream-batch=1, synthetic RGB frame feed (BASELINE.json configs[1])"),
    "sd15-lcm4-512": dict(model="lykon/dreamshaper-8", t=[18, 26, 35, 45], hw=512, gflop=3476.9,
                          metric="frames/sec at 512x512 SD-1.5 + LCM 4-step img2img (stream-batch 4)",
                          name="SD-1.5 + LCM-LoRA 4-step img2img 512x512, stream-batch=4 (BASELINE.json configs[2])"),
    "sd15-lcm4-768": dict(model="lykon/dreamshaper-8", t=[18, 26, 35, 45], hw=768, gflop=9185.6,
                          metric="frames/sec at 768x768 SD-1.5 + LCM 4-step img2img (stream-batch 4)",
                          name="SD-1.5 + LCM-LoRA 4-step img2img 768x768, stream-batch=4, synthetic feed (BASELINE.json configs[4] without codecs)"),
}
MODEL_ID = "stabilityai/sd-turbo"
T_INDEX_LIST = [32]
H = W = 512
METRIC = WORKLOADS["sd-turbo-512"]["metric"]
GFLOP_PER_FRAME = 1068.0  # 804.3 (UNet) + 122.3 (TAESD enc) + 141.4 (TAESD dec)
WORKLOAD = WORKLOADS["sd-turbo-512"]["name"]


def select_workload(key: str) -> None:
    global MODEL_ID, T_INDEX_LIST, H, W, METRIC, GFLOP_PER_FRAME, WORKLOAD
    w = WORKLOADS[key]
    MODEL_ID, T_INDEX_LIST, METRIC, GFLOP_PER_FRAME, WORKLOAD = w["model"], w["t"], w["metric"], w["gflop"], w["name"]
    H = W = w["hw"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_uuid: str):
        self.lines = []
        self.proc = None
        self.uuid = gpu_uuid
        self.t = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.uuid, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def run_oracle(steps: int, warmup: int, budget_s: float):
    """Times the fp32 oracle (the reference's CPU diffusers path restated, oracle/) on the host cores."""
    import torch
    from oracle import pipeline as opipe
    from oracle import stream as ostream
    from oracle import unet as ounet
    from oracle import weights as ow
    threads = torch.get_num_threads()   # torch's default = physical cores; oversubscribing the SMT siblings is slower
    cfg = ounet.config_for(MODEL_ID)
    usd = ow.to_float(ow.make_unet_weights(cfg))
    vsd = ow.to_float(ow.make_taesd_weights())
    orc = ostream.StreamOracle(usd, cfg, vsd, T_INDEX_LIST, W, H)
    orc.prepare(ow.make_prompt_embeds(cfg.cross_attention_dim).float(), guidance_scale=0.0)
    frames = [ow.make_frame(H, W, seed=i) for i in range(4)]
    t0 = time.perf_counter()
    opipe.frame_to_u8(orc, frames[0])  # at least one warm-up frame
    first = time.perf_counter() - t0
    w_done = 1
    while w_done < warmup and (time.perf_counter() - t0) + first < budget_s * 0.25:
        opipe.frame_to_u8(orc, frames[w_done % 4])
        w_done += 1
    times = []
    t_start = time.perf_counter()
    for i in range(steps):
        t1 = time.perf_counter()
        opipe.frame_to_u8(orc, frames[i % 4])
        times.append(time.perf_counter() - t1)
        if time.perf_counter() - t_start + times[-1] > budget_s:
            break
    total = sum(times)
    return {"fps": len(times) / total, "ms_per_step": 1000.0 * total / len(times), "steps": len(times), "warmup": w_done,
            "threads": threads}


def main_reference(args):
    rank = int(os.getenv("RANK", "0"))
    if rank != 0:
        return 0
    r = run_oracle(args.steps, args.warmup, budget_s=150.0)
    cb = {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "port",
          "sample": f"{r['steps']} full 512x512 SD-Turbo 1-step frames (UNet + TAESD enc/dec, fp32 torch oracle; the reference's "
                    "diffusers/StreamDiffusion packages are not installable offline), time-bounded to 150 s"}
    line = {"impl": "reference", "metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "t_index_list": T_INDEX_LIST, "weights": "seeded synthetic"},
            "cpu_baseline": cb,
            "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def main_gpu(args):
    import torch
    import torch.distributed as dist
    os.environ.setdefault("B200SD_SYNTHETIC_WEIGHTS", "1")
    os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the single JSON line (NCCL prints its version banner otherwise)
    os.environ["NVENC"] = "1"  # keep the output tensor in HBM (lib/pipeline.py:83,96)
    from ai_rtc_agent_b200.host import dist as bdist
    from ai_rtc_agent_b200.host.pipeline import StreamDiffusionPipeline
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference)")
    rank, world, local = bdist.init()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    bdist.load_and_broadcast(MODEL_ID, dev)          # rank 0 materialises, NCCL broadcast, once
    pipe = StreamDiffusionPipeline(MODEL_ID, t_index_list=T_INDEX_LIST, width=W, height=H)
    stream = pipe.model.stream

    g = torch.Generator().manual_seed(1000 + rank)
    ring_host = [torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(64)]
    ring_dev = [f.to(dev) for f in ring_host]
    out_host = torch.empty((1, 3, H, W), dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for i in range(warm):
        pipe(ring_dev[i % 64])
    # ---- device-resident throughput (value)
    sampler = ClockSampler("GPU-" + str(torch.cuda.get_device_properties(dev).uuid)) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pipe(ring_dev[(warm + i) % 64])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    barrier()
    clocks = sampler.stop() if sampler else None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    value = world * args.steps / (ms_total / 1000.0)
    # ---- end to end through the public call with host buffers (e2e)
    lat = []
    barrier()
    t_all = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        frame = ring_host[(warm + i) % 64].to(dev, non_blocking=True)
        out = pipe(frame)
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lat.append(time.perf_counter() - t0)
    e2e_s = torch.tensor([time.perf_counter() - t_all], device=dev)
    barrier()
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_fps = world * args.steps / e2e_s.item()
    p50 = torch.tensor([statistics.median(lat) * 1000.0], device=dev)
    if world > 1:
        dist.all_reduce(p50, op=dist.ReduceOp.MAX)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    # ---- roofline of the dominant kernel: per-launch device times from an eager replay outside the timed region
    prof = stream.profile(ring_dev[0], iters=3)
    by = {}
    for op in prof:
        kind = op["name"].split(" ")[0]
        d = by.setdefault(kind, {"ms": 0.0, "flops": 0.0, "launches": 0})
        d["ms"] += op["ms"]; d["flops"] += op["flops"]; d["launches"] += 1
    peaks, peak_src = measured_peaks()
    # dominant kernel: its launches alone, replayed from their own CUDA graph (same order, buffers, PDL edges and weight
    # streaming as inside the frame graph) -> average launch duration without the host-side gaps of the eager replay
    ig = stream.profile_kind("igemm", iters=20)
    ig_tflops = ig["flops"] / (ig["ms"] * 1e-3) / 1e12
    step_ms = ms_total / args.steps
    step_tflops = GFLOP_PER_FRAME * (value / world) / 1e3
    traffic, traffic_note = None, "no ncu capture found under profiles/"
    tpath = os.path.join(ROOT, "profiles", "igemm_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_note = tj.get("dram_bytes_per_launch"), tj.get("note", "")
    roofline = {
        "bound": "tensor", "kernel": "igemm_kernel (tcgen05 implicit-GEMM conv/linear)",
        "achieved": ig_tflops, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": ig_tflops / peaks["bf16_tflops_sustained"], "traffic": traffic,
        "peak_source": peak_src + ", sustained figure (kernel timed inside a long step)",
        "traffic_note": traffic_note,
        "kernel_share_of_step": ig["ms"] / step_ms, "kernel_launches_per_step": ig["launches"],
        "kernel_ms_per_step": ig["ms"], "kernel_avg_launch_us": 1e3 * ig["ms"] / max(ig["launches"], 1),
        "kernel_timing": "CUDA events around a graph holding only the igemm launches of one frame, 20 replays",
        "kernel_algorithmic_gflop_per_step": ig["flops"] / 1e9,
        "step_achieved": step_tflops, "step_frac": step_tflops / peaks["bf16_tflops_sustained"],
        "step_algorithmic_gflop": GFLOP_PER_FRAME,
        "by_kernel_eager_ms": {k: round(v["ms"], 4) for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])},
    }
    # ---- CPU baseline (reported, not the target): bounded sample on this box's host cores
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = run_oracle(steps=2, warmup=1, budget_s=40.0)
        cpu = {"value": r["fps"], "unit": "frames/s", "cores": r["threads"], "kind": "port",
               "sample": f"{r['steps']} full 512x512 SD-Turbo 1-step frames of the fp32 torch oracle (1 warm-up), same workload"}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "t_index_list": T_INDEX_LIST, "weights": "seeded synthetic (no checkpoint offline)",
                   "parallelism": f"dp{world}: one independent stream per GPU, NCCL weight broadcast at init only",
                   "l2": "UNet weights (1.73 GB) are re-streamed from HBM every step (>> 126 MB L2); 64-frame input ring",
                   "model": MODEL_ID},
        "p50_ms": p50.item(),
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": H * W * 3, "d2h_bytes_per_step": H * W * 3,
                "p50_ms": p50.item()},
        "gpu_launches": stream.launches_per_step * args.steps,
        "launches_per_step": stream.launches_per_step,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="sd-turbo-512", choices=sorted(WORKLOADS))
    a = ap.parse_args()
    select_workload(a.workload)
    sys.exit(main_reference(a) if a.impl == "reference" else main_gpu(a))

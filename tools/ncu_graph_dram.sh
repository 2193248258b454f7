#!/bin/bash
# Whole-frame DRAM traffic: the frame's CUDA graph profiled as ONE workload (ncu --graph-profiling graph), caches NOT flushed
# between its kernels (--cache-control none), so activations that stay L2-resident inside a frame are not counted as DRAM
# traffic the way the per-kernel launch list counts them.  usage: bash tools/ncu_graph_dram.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
B200SD_LANES=1 ncu --graph-profiling graph --cache-control none --clock-control none \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -c 2600 --csv \
    --log-file gpurun_out/graph_dram_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline \
    > gpurun_out/ncu_graph_${TAG}.log 2>&1
python - "$TAG" <<'PY'
import csv, sys, collections
tag = sys.argv[1]
rows = collections.OrderedDict()
for r in csv.DictReader(l for l in open(f"gpurun_out/graph_dram_{tag}.csv") if l.startswith('"')):
    d = rows.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
big = [d for d in rows.values() if d.get("gpu__time_duration.sum", 0) > 2.0e6]   # > 2 ms: the frame graphs
print(f"{len(rows)} profiled workloads, {len(big)} frame graphs")
for d in big[-4:]:
    print(f"  {d['name'][:40]:40s} {d['gpu__time_duration.sum']/1e6:.3f} ms  DRAM read {d['dram__bytes_read.sum']/1e9:.3f} GB  write {d['dram__bytes_write.sum']/1e9:.3f} GB")
PY

#!/usr/bin/env python
"""Markdown summary of an `ncu --set full` report: python tools/ncu_full_summary.py <file.ncu-rep> <title> > profiles/<name>.md
Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps the metrics the roofline discussion uses."""
import csv
import io
import subprocess
import sys

KEEP = [
    "launch__grid_size", "launch__block_size", "launch__cluster_size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__inst_executed.sum", "sm__sass_inst_executed_op_local_st.sum", "sm__sass_inst_executed_op_local_ld.sum",
]


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    print(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`; report: {rep} (not committed).  "
          "Times are cold-cache, serialised launches.\n")
    seen = {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kname = r[col["Kernel Name"]]
        key = (kname, r[col.get("Grid Size", col["Kernel Name"])])
        seen[key] = seen.get(key, 0) + 1
        if seen[key] > 1:
            continue   # first instance of each (kernel, grid)
        print(f"## {kname}  grid {r[col['Grid Size']]}  block {r[col['Block Size']]}")
        for m in KEEP:
            if m in col and r[col[m]] != "":
                print(f"- `{m}` = {r[col[m]]} {units[col[m]]}")
        print()


if __name__ == "__main__":
    main()

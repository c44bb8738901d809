#!/usr/bin/env python
"""Summarise an .ncu-rep (full set) into one CSV row per launch: duration, DRAM bytes, tensor pipe, occupancy, stalls.
Usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [out.csv]"""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
]

def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in WANT if c in idx]
    extra = [h for h in hdr if ("tensor" in h and h not in cols)][:6]
    cols += extra
    w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
    w.writerow(["id", "kernel", "grid", "block"] + [f"{c} [{units[idx[c]]}]" for c in cols])
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]]
        w.writerow([r[idx["ID"]], name[:90], r[idx["Grid Size"]], r[idx["Block Size"]]] + [r[idx[c]] for c in cols])

if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Fixed metric table from one or more .ncu-rep files (first kernel of each).
usage: ncu_table.py label=path.ncu-rep ...   (needs `ncu` on PATH)"""
import csv
import io
import subprocess
import sys

METRICS = [
    "launch__grid_size", "launch__cluster_dim_x", "launch__registers_per_thread", "gpu__time_duration.sum",
    "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units, v = rows[0], rows[1], rows[2]
    return dict(zip(h, v)), dict(zip(h, units))


def main():
    cols = [a.rsplit("=", 1) for a in sys.argv[1:]]
    data = [(lab, *load(p)) for lab, p in cols]
    print("| metric | unit | " + " | ".join(lab for lab, _, _ in data) + " |")
    print("|---|---|" + "---|" * len(data))
    print("| kernel | | " + " | ".join(d.get("Kernel Name", "")[:34] for _, d, _ in data) + " |")
    for m in METRICS:
        if not any(m in d for _, d, _ in data):
            continue
        u = next((un.get(m, "") for _, d, un in data if m in d), "")
        print("| %s | %s | " % (m, u) + " | ".join(d.get(m, "") for _, d, _ in data) + " |")


if __name__ == "__main__":
    main()

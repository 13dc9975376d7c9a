#!/usr/bin/env python
"""DRAM traffic of the tensor-core conv launches of one SSD300 step from an ncu CSV
(--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv_tc_kernel).
usage: conv_traffic.py conv_traffic.csv launches_per_step out.json"""
import csv
import json
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    path, per_step, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = next(rd)
    iid, im, iu, iv = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    launches = {}
    for r in rd:
        if len(r) <= iv:
            continue
        d = launches.setdefault(int(r[iid]), {})
        if r[im].startswith("dram__bytes"):
            d[r[im]] = to_bytes(r[iv], r[iu])
        elif r[im] == "gpu__time_duration.sum":
            v = float(r[iv].replace(",", ""))
            d["us"] = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    ids = sorted(launches)
    n_groups = len(ids) // per_step
    assert n_groups >= 1, "not enough launches captured"
    grp = ids[(n_groups - 1) * per_step:n_groups * per_step]   # the last complete step
    rd_b = sum(launches[i].get("dram__bytes_read.sum", 0) for i in grp)
    wr_b = sum(launches[i].get("dram__bytes_write.sum", 0) for i in grp)
    res = {"launches_per_step": per_step, "dram_read_bytes_per_step": rd_b, "dram_write_bytes_per_step": wr_b,
           "dram_bytes_per_launch_avg": (rd_b + wr_b) / per_step,
           "ncu_time_us_per_step": sum(launches[i].get("us", 0) for i in grp),
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:conv_tc_kernel on bench.py, last complete step"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

"""Summarise ncu captures for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py full  gpurun_out/x.ncu-rep [...] > profiles/rNN_ncu_full_<kernel>.json
    python tools/ncu_summary.py shares gpurun_out/launches.csv   > profiles/rNN_launch_shares.txt
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__cluster_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second",
]


def kernel_name(full):
    m = re.search(r"(\w+_kernel)", full) if "ehb::" in full else None
    return "ehb::" + m.group(1) if m else re.sub(r"[<(].*", "", full)[:60]


def full(paths):
    out = []
    for p in paths:
        raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        head, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = {"capture": p.split("/")[-1], "kernel": kernel_name(r[head.index("Kernel Name")])}
            for k in KEEP:
                if k in head:
                    i = head.index(k)
                    rec[k] = f"{r[i]} {units[i]}".strip()
            out.append(rec)
    print(json.dumps(out, indent=1))


def shares(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = {}, collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1000.0 if row["Metric Unit"].startswith("n") else v
        k = kernel_name(row["Kernel Name"])
        tot[k] = tot.get(k, 0.0) + us
        cnt[k] += 1
    T = sum(tot.values())
    print(f"one sampling pass (cfg 2: 64 images x 10 samples, DDIM-5), ncu --metrics gpu__time_duration.sum, "
          f"serialised/cold-cache: total {T:.0f} us over {sum(cnt.values())} launches")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:24]:
        print(f"{v:9.1f} us {100 * v / T:5.1f}% x{cnt[k]:3d} {k}")


if __name__ == "__main__":
    (full if sys.argv[1] == "full" else lambda a: shares(a[0]))(sys.argv[2:])

"""Summarise ncu output into small, committed text files.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r01.csv > profiles/r01_launches.md
    python profiles/summarize_ncu.py full gpurun_out/prof_r01.ncu-rep   > profiles/r01_full.md
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    print("| kernel | launches | avg us | share of GPU time |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    print(f"\ntotal launches {sum(len(v) for v in agg.values())}, total GPU time {tot / 1e6:.3f} ms "
          "(ncu per-launch times are cold-cache and serialised: compare shares, not absolutes)")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"\n### `{r[h.index('Kernel Name')][:100]}`  (launch id {r[0]})\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in h:
                print(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |")
        rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        print(f"| traffic (dram read+write, as reported) | {r[rd]} + {r[wr]} | {units[rd]} |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

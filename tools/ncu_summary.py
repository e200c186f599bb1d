#!/usr/bin/env python3
"""Condenses ncu output into the small text files kept under profiles/.

    python tools/ncu_summary.py report  gpurun_out/prof.ncu-rep  profiles/r01_xyz.csv     # key metrics per profiled launch
    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_summary.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def report(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow([f"{k} [{units[idx[k]]}]" if units[idx[k]] else k for k in keys])
        for r in rows[2:]:
            w.writerow([r[idx[k]] for k in keys])
    print(open(out).read())


def launches(src, out):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        v = v / 1000 if r[i_unit] == "ns" else (v * 1000 if r[i_unit] == "ms" else v)
        n = r[i_name].split("(")[0]
        tot[n] += v
        cnt[n] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {sum(cnt.values())} launches, {T/1000:.2f} ms total\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        for n in sorted(tot, key=lambda k: -tot[k]):
            f.write(f"{n:60s} launches {cnt[n]:5d}  total {tot[n]:12.1f} us  avg {tot[n]/cnt[n]:9.1f} us  share {100*tot[n]/T:5.1f}%\n")
    print(open(out).read())


if __name__ == "__main__":
    {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])

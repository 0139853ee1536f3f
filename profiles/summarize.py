"""Summarise an ncu launch list (gpu__time_duration.sum CSV) and a --set full report into markdown.
usage: python profiles/summarize.py <launches.csv> [<prof.ncu-rep>]"""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        k = row["Kernel Name"].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / tot:.3f} |")


WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [i for i, h in enumerate(hdr) if h in WANT]
    print("\n| metric | unit | " + " | ".join(f"launch {k}" for k in range(len(rows) - 2)) + " |")
    print("|---|---|" + "---|" * (len(rows) - 2))
    for i in idx:
        print(f"| {hdr[i]} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")


if __name__ == "__main__":
    launches(sys.argv[1])
    if len(sys.argv) > 2:
        full(sys.argv[2])

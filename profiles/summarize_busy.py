"""Per-launch table of one objective evaluation ("round") of one slot group from the csv written by run_ncu3.sh.
usage: python profiles/summarize_busy.py <busy.csv> <tag>"""
import collections
import csv
import sys

path, tag = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("=="))]
h = rows[0]
ki, gi, vi, mi, ii, si = (h.index(k) for k in ("Kernel Name", "Grid Size", "Metric Value", "Metric Name", "ID", "Stream"))
L = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    d = L.setdefault(r[ii], {"k": r[ki].split("(")[0].replace("void ", ""), "g": r[gi], "s": r[si]})
    d[r[mi]] = float(r[vi].replace(",", ""))
items = list(L.values())
one = [x for x in items if x["s"] == items[0]["s"]]
i0 = [i for i, x in enumerate(one) if x["k"].startswith("k_build")][0]
B = "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_elapsed"
print(f"# {tag}: one objective evaluation (\"round\") of one slot group, launch by launch\n")
script, cmd = ("profiles/capture_r02.sh (pass 2)", "python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e") \
    if tag.startswith("r02") else ("profiles/run_ncu3.sh", "python bench.py --experts-per-step 592 --steps 1 --warmup 1")
print(f"`bash {script}` = `ncu --metrics gpu__time_duration.sum," + B + ",dram__bytes_read.sum,"
      f"dram__bytes_write.sum --clock-control none` on `{cmd}` "
      "(c3 workload; launches serialised by ncu, cold caches).\n")
print("| kernel | grid | time (us) | DMMA pipe busy (% of elapsed) | DRAM read+write (MB) |\n|---|---|---|---|---|")
tot = collections.OrderedDict()
for x in one[i0:]:
    if x["k"].startswith("k_build") and x is not one[i0]:
        break
    t, b = x["gpu__time_duration.sum"] / 1e3, x[B]
    mb = (x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"]) / 1e6
    print(f"| {x['k']} | {x['g']} | {t:.1f} | {b:.1f} | {mb:.0f} |")
    a = tot.setdefault(x["k"], [0, 0, 0, 0])
    a[0] += t
    a[1] += t * b / 100
    a[2] += mb
    a[3] += 1
T = sum(a[0] for a in tot.values())
print("\n| kernel | launches | total us | share of round | time-weighted DMMA busy | DRAM MB |\n|---|---|---|---|---|---|")
for k, a in tot.items():
    print(f"| {k} | {a[3]} | {a[0]:.1f} | {a[0] / T:.3f} | {a[1] / a[0]:.3f} | {a[2]:.0f} |")
print(f"\nRound total {T / 1e3:.2f} ms.")

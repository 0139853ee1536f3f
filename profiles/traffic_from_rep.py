"""profiles/r02_traffic.json (the source of bench.py's roofline.traffic) from the captures of profiles/capture_r02.sh:
the ncu --set full report (three middle panel launches, read here with `ncu -i`) and, for the average over EVERY panel
launch of one round, the per-launch dram__bytes of pass 2 (r02_busy.csv).
usage: python profiles/traffic_from_rep.py gpurun_out/prof_r02_potrf.ncu-rep potrf 1024 c3 [gpurun_out/r02_busy.csv]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, phase, eps, workload = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units, body = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(h)}


def val(r, name):
    v = float(r[col[name]].replace(",", ""))
    u = units[col[name]].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9,
                "msecond": 1e-3, "usecond": 1e-6, "second": 1.0, "nsecond": 1e-9}.get(u, 1.0)


launches = []
for r in body:
    launches.append({"grid": r[col["Grid Size"]], "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                     "seconds": val(r, "gpu__time_duration.sum"),
                     "dmma_busy_pct_active": float(r[col["sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"]])})
out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_traffic.json")
try:
    out = json.load(open(out_path))
except (OSError, ValueError):
    out = {}
out.update(workload=workload, experts_per_step=eps)
full = {"launches": len(launches), "dram_bytes_per_launch": sum(x["dram_bytes"] for x in launches) / len(launches),
        "seconds_per_launch_under_ncu": sum(x["seconds"] for x in launches) / len(launches),
        "dmma_busy_pct_of_active": [x["dmma_busy_pct_active"] for x in launches],
        "grids": [x["grid"] for x in launches],
        "source": os.path.basename(rep) + " (ncu --set full --clock-control none --import-source on, three consecutive "
                  "middle panel launches of one round of one slot group)"}
out[phase] = {"kernel": body[0][col["Kernel Name"]], "full_set_capture": full}
if len(sys.argv) > 5:      # every launch of the kernel in one round of one slot group, from the light pass
    rows2 = [r for r in csv.reader(l for l in open(sys.argv[5]) if not l.startswith("=="))]
    h2 = rows2[0]
    ki, vi, mi, ii, si, gi = (h2.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Stream", "Grid Size"))
    per = {}
    for r in rows2[1:]:
        if len(r) > vi:
            per.setdefault(r[ii], {"k": r[ki], "s": r[si], "g": r[gi]})[r[mi]] = float(r[vi].replace(",", ""))
    items = list(per.values())
    one = [x for x in items if x["s"] == items[0]["s"]]
    i0 = [i for i, x in enumerate(one) if "k_build" in x["k"]][0]
    rnd = []
    for x in one[i0 + 1:]:
        if "k_build" in x["k"]:
            break
        rnd.append(x)
    mine = [x for x in rnd if body[0][col["Kernel Name"]].split("(")[0] in x["k"]]
    tot = sum(x["dram__bytes_read.sum"] + x["dram__bytes_write.sum"] for x in mine)
    out[phase].update(launches=len(mine), dram_bytes_per_launch=tot / len(mine), dram_bytes_per_round=tot,
                      grids=[x["g"] for x in mine],
                      source=os.path.basename(sys.argv[5]) + " (ncu dram__bytes_read.sum + dram__bytes_write.sum of every "
                             "launch of the kernel in one round of one slot group; the full-set capture of three of "
                             "them is under full_set_capture)")
else:
    out[phase].update(launches=full["launches"], dram_bytes_per_launch=full["dram_bytes_per_launch"],
                      source=full["source"])
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out[phase], indent=1))

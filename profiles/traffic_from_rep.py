"""profiles/r02_traffic.json from the ncu --set full capture of profiles/capture_r02.sh (read here with `ncu -i`).
usage: python profiles/traffic_from_rep.py gpurun_out/prof_r02_potrf.ncu-rep potrf 1024 c3"""
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
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0,
                "nsecond": 1e-9}.get(u, 1.0)


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
out[phase] = {"kernel": body[0][col["Kernel Name"]], "launches": len(launches),
              "dram_bytes_per_launch": sum(x["dram_bytes"] for x in launches) / len(launches),
              "dram_bytes_total": sum(x["dram_bytes"] for x in launches),
              "seconds_total_under_ncu": sum(x["seconds"] for x in launches),
              "grids": [x["grid"] for x in launches],
              "source": os.path.basename(rep) + " (ncu --set full --clock-control none, consecutive launches of one "
                        "round of one slot group)"}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out[phase], indent=1))

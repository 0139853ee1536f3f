"""Summarise a GPSAT_TRACE file (api.cu, gpsat_gpr_optimise): per optimiser call, the share of rounds that were queued
while the group's previous round was still running on the device (dry = 0), and the host-side interval between
consecutive queueings of a group.  usage: python profiles/trace_summary.py <trace file> [...]"""
import sys

import numpy as np

for path in sys.argv[1:]:
    calls, cur = [], None
    for ln in open(path):
        if ln.startswith("#"):
            cur = {"hdr": ln[2:].strip(), "rows": []}
            calls.append(cur)
        elif cur is not None:
            f = ln.strip().split(",")
            if len(f) >= 7:
                cur["rows"].append([float(x) for x in f[:7]])
    for c in calls:
        r = np.array(c["rows"])
        if len(r) < 10:
            continue
        live = r[r[:, 3] > 0]
        later = live[live[:, 0] > 0]
        dry = later[:, 6].mean() if len(later) else float("nan")
        gaps = []
        for g in np.unique(live[:, 1]):
            t = live[live[:, 1] == g][:, 2]
            gaps.append(np.diff(t))
        gaps = np.concatenate(gaps) if gaps else np.zeros(0)
        print(f"{path}: {c['hdr']}: {len(live)} rounds queued; queued while the previous round was still running: "
              f"{100 * (1 - dry):.1f} %; host interval between rounds of a group: median {np.median(gaps) / 1e3:.2f} ms, "
              f"p99 {np.percentile(gaps, 99) / 1e3:.2f} ms")

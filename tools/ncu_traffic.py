#!/usr/bin/env python
"""DRAM traffic per frame of every kernel of an ncu --set full capture -> profiles/ncu_traffic.json (read by bench.py for
`roofline.traffic`).  Usage: ncu_traffic.py frames_per_launch file.ncu-rep [more.ncu-rep ...]   (later reports fill in
kernels the earlier ones did not capture)"""
import csv
import json
import os
import subprocess
import sys

STAGE = {"k_reproject_hoist4": "reproject", "k_reproject_hoist": "reproject", "k_grad_circ": "grad_circ", "k_grad_rowscan": "grad_rowscan", "k_colscan": "colscan",
         "k_circ_stream_rs": "circ_peaks", "k_peaks_emit": "peaks_emit", "k_sat_check_g": "sat_check", "k_sat_check_rs": "sat_check", "k_fallback_frame": "sat_check",
         "k_peaks_prepare": "prepare"}
frames, reps = float(sys.argv[1]), sys.argv[2:]
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
try:
    sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(dst)).stdout.strip()
except Exception:  # noqa: BLE001
    sha = "?"
res = {"sources": [os.path.basename(r) for r in reps], "source": "ncu --set full of tools/prof_step.py, " + ", ".join(os.path.basename(r) for r in reps) + ", tree " + sha,
       "frames_per_launch": frames, "kernels": {}}
h, units = [], []


def col(r, name):
    i = h.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


for rep in reps:
  out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  h, units = rows[0], rows[1]
  seen = set(res["kernels"])
  for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    for k, st in STAGE.items():
        if (k + "(" in name or k + "<" in name) and (st not in seen or st == "sat_check"):
            rd, wr = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
            e = res["kernels"].setdefault(st, {"kernel": k, "dram_bytes_per_frame": 0.0, "us_per_frame_under_ncu": 0.0})
            e["dram_bytes_per_frame"] += (rd + wr) / frames
            i = h.index("gpu__time_duration.sum")
            e["us_per_frame_under_ncu"] += float(r[i].replace(",", "")) * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "ns": 1e-3, "nsecond": 1e-3}.get(units[i].lower(), 1) / frames
json.dump(res, open(dst, "w"), indent=1)
print(json.dumps(res, indent=1))

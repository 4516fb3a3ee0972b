#!/usr/bin/env python
"""profiles/ncu_traffic.json entry from an .ncu-rep:  python tools/ncu_traffic.py <kernel key, e.g. spmm_c256> <rep> [note]"""
import csv, json, os, subprocess, sys
key, rep = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(vals, units)))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = sum(float(d[m][0]) * scale[d[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", "ncu_traffic.json")
cur = json.load(open(path)) if os.path.exists(path) else {}
cur[key] = {"dram_bytes": tot, "kernel_name": d["Kernel Name"][0], "duration_us_under_ncu": float(d["gpu__time_duration.sum"][0]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}[d["gpu__time_duration.sum"][1]],
            "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} {note}".strip()}
json.dump(cur, open(path, "w"), indent=1, sort_keys=True)
print(key, cur[key])

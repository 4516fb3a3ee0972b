#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv):  python tools/launch_list_summary.py launches.csv "<title>" """
import collections
import csv
import sys

path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
tot, cnt = collections.Counter(), collections.Counter()
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
    name = r[ix["Kernel Name"]]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print(title)
print("gpu__time_duration.sum per kernel, cold-cache and serialised under ncu: compare SHARES with bench.py kernel_families, not absolutes.")
print(f"total {total:.2f} ms over {sum(cnt.values())} launches\n")
for name, ms in tot.most_common():
    print(f"{ms:9.3f} ms {100 * ms / total:5.1f} %  x{cnt[name]:4d}  {name[:110]}")

#!/usr/bin/env python
"""Per-opcode executed-instruction totals and the hottest SASS lines of an .ncu-rep (source page):
   python tools/ncu_hot.py rep.ncu-rep [top_n]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = collections.Counter(); n = 0
for r in rows:
    try: ex = int(r["Instructions Executed"])
    except (ValueError, KeyError): continue
    ins = r["Source"].split()
    op = ins[1] if ins and ins[0].startswith("@") else (ins[0] if ins else "?")
    tot[op.split(".")[0]] += ex; n += ex
print("total warp instructions", n)
print("  ".join(f"{k}={v/n*100:.1f}%" for k, v in tot.most_common(24)))
print("hottest by stall samples:")
rows2 = sorted(rows, key=lambda r: -int(r.get("# Samples") or 0))[:top]
for r in rows2:
    st = {k[6:]: int(v) for k, v in r.items() if k.startswith("stall_") and "(Not" not in k and v and v.isdigit() and int(v) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{r["# Samples"]:>7} {r["Instructions Executed"]:>10}  {r["Source"].strip()[:70]:70s} {st}')

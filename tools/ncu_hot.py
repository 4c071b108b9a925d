#!/usr/bin/env python
"""Hottest SASS instructions of a kernel in an .ncu-rep (stall samples):  python tools/ncu_hot.py rep kernel_regex [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
iS, iSamp, iEx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iSamp] or 0) for r in body)
print("instructions", len(body), "samples", tot, "warp-inst executed", sum(int(r[iEx] or 0) for r in body))
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda k: -int(body[k][iSamp] or 0))[:top]
for k in sorted(order):
    r = body[k]
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[iSamp]):6d} {100*int(r[iSamp])/max(tot,1):5.1f}% ex={r[iEx]:>8s} {r[iS].strip()[:70]:70s} {st}")

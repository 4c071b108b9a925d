#!/usr/bin/env python
"""Warp instructions executed and stall samples per SOURCE LINE of a kernel in an .ncu-rep (built with -lineinfo, captured with
--import-source on):  python tools/ncu_lines.py rep [kernel_regex] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
rx = sys.argv[2] if len(sys.argv) > 2 else "."
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[start]
iEx, iSamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
per = {}
cur = None
for r in rows[start + 1:]:
    if not r or r[0] in ("File Path", "Function Name", "Line No"):
        if r and r[0] == "Function Name" and per:
            break
        continue
    if r[0] and len(r) > iEx and r[iEx] not in ("", "-"):   # a line's row carries the totals of its SASS rows
        cur = (int(r[0]), r[1].strip())
        e = per.setdefault(cur, [0, 0])
        e[0] += int(r[iEx])
        e[1] += int(r[iSamp] or 0)
tot = sum(v[0] for v in per.values()) or 1
ts = sum(v[1] for v in per.values()) or 1
print("warp instructions", tot, "samples", ts)
for (ln, src), (ex, sm) in sorted(sorted(per.items(), key=lambda kv: -kv[1][0])[:top]):
    print(f"{ln:5d} {100 * ex / tot:5.1f}% inst {100 * sm / ts:5.1f}% samp  {src[:110]}")

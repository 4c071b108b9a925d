#!/usr/bin/env python
"""Condense an `ncu --csv` launch list (long format: one row per launch and metric) into one row per launch of OUR
kernels plus a per-kernel summary, and derive the dominant kernel's DRAM traffic per launch.
    python tools/launch_list.py raw.csv out.csv traffic.json"""
import csv
import json
import sys
from collections import OrderedDict, defaultdict

raw, out, traffic_path = sys.argv[1:4]
OURS = ("scan_", "cross_", "stft", "ss2d_", "map_")
rows = []
with open(raw, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
launches = OrderedDict()
total_all = 0.0
for r in rd:
    try:
        key = int(r["ID"])
        val = float(r["Metric Value"].replace(",", ""))
    except (KeyError, ValueError):
        continue
    unit = r.get("Metric Unit", "")
    name = r["Metric Name"]
    if name == "gpu__time_duration.sum":
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)  # -> us
    elif name.startswith("dram__bytes"):
        val *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    dims = lambda t: "x".join(v.strip() for v in t.strip("()").split(","))
    e = launches.setdefault(key, {"kernel": r["Kernel Name"], "grid": dims(r.get("Grid Size", "")), "block": dims(r.get("Block Size", ""))})
    e[name] = val
ours = [(k, e) for k, e in launches.items() if any(t in e["kernel"] for t in OURS)]
t_all = sum(e.get("gpu__time_duration.sum", 0.0) for e in launches.values())
t_ours = sum(e.get("gpu__time_duration.sum", 0.0) for _, e in ours)
per = defaultdict(lambda: [0, 0.0, 0.0])
with open(out, "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-core --no-stft --no-cpu-baseline` (our kernels only; "
            "cold-cache, serialised: compare shares, not absolutes)\n")
    f.write("id,kernel,grid,block,time_us,dram_read_bytes,dram_write_bytes\n")
    for k, e in ours:
        name = e["kernel"].split("(")[0].replace(", ", ";").replace(",", ";")
        t, rb, wb = e.get("gpu__time_duration.sum", 0.0), e.get("dram__bytes_read.sum", 0.0), e.get("dram__bytes_write.sum", 0.0)
        f.write(f"{k},{name},{e['grid']},{e['block']},{t:.2f},{rb:.0f},{wb:.0f}\n")
        p = per[name]
        p[0] += 1
        p[1] += t
        p[2] += rb + wb
    f.write("# per-kernel summary: kernel,launches,total_us,share_of_our_kernels,avg_us,avg_dram_bytes\n")
    for name, (n, t, by) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        f.write(f"# {name},{n},{t:.1f},{t / max(t_ours, 1e-9):.3f},{t / n:.2f},{by / n:.0f}\n")
    f.write(f"# all launches profiled: {len(launches)}, total {t_all:.1f} us; ours: {len(ours)}, total {t_ours:.1f} us\n")
# dominant kernel = the multi-chunk backward kernel (every backward call with seqlen > 2048)
dom = [e for _, e in ours if "scan_bwd_pipe_kernel" in e["kernel"]]
if dom:
    tr = sum(e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0) for e in dom) / len(dom)
    json.dump({"traffic_bytes_per_launch": round(tr), "launches_averaged": len(dom),
               "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over the bench command's launches of the dominant kernel"},
              open(traffic_path, "w"))
print(f"launch list: {len(launches)} launches, {len(ours)} ours, dominant {len(dom)}")

import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
body = []
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
iS, iEx = hdr.index("Source"), hdr.index("Instructions Executed")
mix = collections.Counter()
tot = 0
for r in body:
    ex = int(r[iEx] or 0)
    src = r[iS].strip()
    toks = src.split()
    op = toks[0]
    if op.startswith("@"):
        op = toks[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "LD", "ST", "MUFU", "SHFL", "RED", "ATOM")) and "." in op else "")
    mix[op] += ex
    tot += ex
elems = float(sys.argv[3]) if len(sys.argv) > 3 else 1
print("total warp-inst", tot, "per element (thread-inst)", tot * 32 / elems)
for op, n in mix.most_common(45):
    print(f"{op:14s} {n:10d} {100*n/tot:5.1f}%  {n*32/elems:6.2f}/elem")

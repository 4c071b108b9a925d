#!/usr/bin/env python
"""Static look at a kernel's loops without a GPU: disassemble an object file, find every backward branch of one kernel and
print the loop body's instruction count, its mix and the source lines it comes from.
    python tools/sass_loops.py vm_asr_b200/lib/obj/scan_bwd_pipe.o 'scan_bwd_pipe_kernelILb1ELi4ELb0' [--min 100]"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel", help="substring of the mangled kernel name")
    ap.add_argument("--min", type=int, default=100)
    ap.add_argument("--lines", action="store_true", help="per source line counts of every loop")
    ap.add_argument("--max", type=int, default=1000)
    ap.add_argument("--dump", type=int, nargs=2, help="print the instructions [lo, hi] with their source lines")
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(args.obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        text = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    sect = None
    instrs = []   # (index, opcode, full text, file, line)
    labels = {}
    cur = ("", 0)
    for ln in text.splitlines():
        if ln.startswith(".text."):
            sect = ln
            continue
        if sect is None or args.kernel not in sect:
            if ln.startswith(".section") or ln.startswith("//----"):
                if sect and args.kernel in sect and instrs:
                    break
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"^(\.L_x?_?\w+):", ln)
        if m:
            labels[m.group(1)] = len(instrs)
            continue
        m = re.match(r"^\s+/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            body = m.group(1)
            toks = body.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            instrs.append((len(instrs), op, body, cur[0], cur[1]))
    print(f"{len(instrs)} instructions")
    if args.dump:
        for i, op, body, f, l in instrs[args.dump[0]:args.dump[1] + 1]:
            print(f"{i:6d} {f}:{l:<4d} {body}")
        return
    loops = []
    for i, op, body, f, l in instrs:
        if op.startswith("BRA"):
            m = re.search(r"`\((\.L_\w+)\)", body)
            if m and m.group(1) in labels and labels[m.group(1)] <= i:
                loops.append((labels[m.group(1)], i))
    for lo, hi in loops:
        n = hi - lo + 1
        if n < args.min or n > args.max:
            continue
        mix = collections.Counter()
        src = collections.Counter()
        for _, op, body, f, l in instrs[lo:hi + 1]:
            mix[op.split(".")[0]] += 1
            src[(f, l)] += 1
        files = collections.Counter()
        for (f, l), c in src.items():
            files[f] += c
        lines_here = [l for (f, l) in src if f.endswith(".cu")]
        print(f"loop [{lo}, {hi}] {n} instructions; {os.path.basename(args.obj)[:-2]}.cu lines {min(lines_here) if lines_here else '-'}..{max(lines_here) if lines_here else '-'}")
        print("   mix:", ", ".join(f"{k} {v}" for k, v in mix.most_common(24)))
        if args.lines:
            for (f, l), c in sorted(src.items()):
                print(f"      {f}:{l} {c}")


if __name__ == "__main__":
    main()

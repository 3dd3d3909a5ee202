#!/usr/bin/env python
"""Instruction histogram of the hottest loop (largest backward branch span) of one kernel in a cubin.

usage: sass_loop_hist.py <cubin> <kernel> [--list]
"""
import collections
import re
import subprocess
import sys


def main():
    cubin, kernel = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    on, ins = False, []
    for line in sass.splitlines():
        if "Function :" in line:
            on = line.split(":")[1].strip() == kernel
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if on and m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = (0, 0, 0)
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr and addr - int(m.group(1), 16) > best[0]:
            best = (addr - int(m.group(1), 16), int(m.group(1), 16), addr)
    _, lo, hi = best
    body = [(a, t) for a, t in ins if lo <= a <= hi]
    ops = collections.Counter()
    for _, t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0]] += 1
    fp64 = sum(ops[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"{kernel}: loop 0x{lo:x}..0x{hi:x}, {len(body)} instructions, {fp64} on the FP64 pipe")
    for k, v in ops.most_common():
        print(f"  {v:4d} {k}")
    if "--list" in sys.argv:
        for a, t in body:
            print(f"/*{a:04x}*/ {t}")


if __name__ == "__main__":
    main()

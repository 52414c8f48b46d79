#!/usr/bin/env python3
"""Summarise an ncu report of the receiver kernel: executed warp-instructions per sample step, and the SASS regions
(contiguous instructions with the same execution count) that hold the instructions and the stall samples.

usage: tools/ncu_regions.py REPORT.ncu-rep WARPS SAMPLES_PER_STREAM
"""
import collections
import csv
import math
import subprocess
import sys


def main():
    rep, warps, samples = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ie, ss = h.index("Instructions Executed"), h.index("# Samples")
    cols = {n: h.index(n) for n in h if n.startswith("stall_") and "Not Issued" not in n}
    data = [r for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    tots = sum(int(r[ss]) for r in data)
    print(f"kernel: {rows[0][1] if len(rows[0]) > 1 else '?'}")
    print(f"executed warp-instructions {tot:,}; per sample step {tot / (warps * samples):.1f}; stall samples {tots:,}")
    regs, cur = [], None
    for i, r in enumerate(data):
        e = int(r[ie])
        key = 0 if e == 0 else round(math.log10(e) * 3)
        if cur is None or key != cur["key"]:
            if cur:
                regs.append(cur)
            cur = {"key": key, "i0": i, "i1": i, "e": 0, "s": 0, "st": collections.Counter(), "ops": collections.Counter()}
        cur["i1"] = i
        cur["e"] += e
        cur["s"] += int(r[ss])
        op = r[1].strip().split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        cur["ops"][op.split(".")[0]] += 1
        for n, c in cols.items():
            if r[c].isdigit():
                cur["st"][n] += int(r[c])
    regs.append(cur)
    print("sass range        n   exec/instr   %inst  %stall  top stalls | top opcodes")
    for g in regs:
        if g["e"] / tot > 0.01 or g["s"] / max(tots, 1) > 0.01:
            n = g["i1"] - g["i0"] + 1
            top = ", ".join(f"{k[6:]}={v / max(g['s'], 1) * 100:.0f}%" for k, v in g["st"].most_common(3))
            ops = ", ".join(f"{k}:{v}" for k, v in g["ops"].most_common(6))
            print(f"[{g['i0']:5d}..{g['i1']:5d}] {n:4d} {g['e'] / n:12.3g} {g['e'] / tot * 100:6.1f} {g['s'] / max(tots, 1) * 100:7.1f}  {top} | {ops}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-source-line stall samples of an ncu report (compiled with -lineinfo, captured with --import-source on): which
lines of which file hold a warp's time.  With --roles the lines of same_rx_pipe_kernel are grouped into the stages of
its four warps, as shares of ONE warp's time (all samples / 4: four warps are resident for the whole kernel).

usage: tools/ncu_source_lines.py REPORT.ncu-rep [--top N] [--roles]
"""
import collections
import csv
import subprocess
import sys


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    hdr, cur, kernel, lines = None, None, None, []
    for r in csv.reader(out.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) == 2 and r[0] == "Function Name":
            kernel = r[1]
        elif len(r) >= 60 and r[0] == "Line No":
            hdr = r
        elif hdr is not None and len(r) >= 60 and r[0].strip().isdigit():
            # the source text may hold commas: count the metric columns from the right
            n = len(hdr)
            i_all = hdr.index("Warp Stall Sampling (All Samples)") - n
            i_inst = hdr.index("Instructions Executed") - n
            try:
                s, ins = int(r[i_all]), int(r[i_inst])
            except ValueError:
                continue
            lines.append((s, ins, cur, int(r[0]), ",".join(r[1:len(r) + i_all - 1]).strip()[:100]))
    return kernel, lines


# stages of same_rx_pipe_kernel by (file, first line, last line): keep in step with same_kernels.cu
PIPE_STAGES = [
    ("same_kernels.cu", 1003, 1020, "matched filter chain, 42 taps (consumer: mark, warp S: space)"),
    ("sm_100_rt.hpp", 0, 10**9, "matched filter chain, 42 taps (consumer: mark, warp S: space)"),
    ("same_lane.cuh", 30, 36, "hypot_fixed (consumer + warp S)"),
    ("same_kernels.cu", 1085, 1131, "warp A: AGC loop + waits"),
    ("same_fast.cuh", 217, 222, "agc_step (warp A; consumer fallback)"),
    ("same_kernels.cu", 1133, 1143, "warp S: loop + waits"),
    ("same_kernels.cu", 740, 800, "warp P: refill loop + waits"),
    ("same_fast.cuh", 0, 216, "warp P: s16 unpack + DC blocker"),
    ("same_kernels.cu", 1154, 1162, "consumer: barrier DONE + exit vote"),
    ("same_kernels.cu", 1163, 1173, "consumer: segment length"),
    ("same_kernels.cu", 1174, 1192, "consumer: AGC fallback"),
    ("same_kernels.cu", 1193, 1208, "consumer: publish position / gain / request"),
    ("same_kernels.cu", 1209, 1217, "consumer: pre-TED + barrier SPACE"),
    ("same_kernels.cu", 1218, 1224, "consumer: soft symbol, TED call, fire"),
    ("same_kernels.cu", 1225, 1240, "consumer: symbol dispatch"),
    ("same_lane.cuh", 37, 70, "consumer: fire_clock"),
    ("same_lane.cuh", 340, 399, "consumer: ted_step"),
    ("same_lane.cuh", 400, 10**9, "consumer: squelch + framer glue"),
    ("same_lane.cuh", 170, 339, "consumer: DFE"),
    ("same_transport.cuh", 0, 10**9, "consumer: framer / transport"),
]


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    kernel, lines = load(rep)
    tot = sum(a[0] for a in lines)
    print(f"kernel: {kernel}")
    print(f"stall samples with a source line: {tot:,}")
    if "--roles" in sys.argv:
        per_warp = tot / 4.0
        acc = collections.OrderedDict()
        for s, ins, f, ln, _ in lines:
            name = next((n for (ff, a, b, n) in PIPE_STAGES if ff == f and a <= ln <= b), "other (" + f + ")")
            e = acc.setdefault(name, [0, 0])
            e[0] += s
            e[1] += ins
        print("\nstage                                                            samples  % of one warp   warp-instructions")
        for name, (s, ins) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
            if s / per_warp >= 0.002:
                print(f"{name:64s} {s:8d} {100 * s / per_warp:10.1f}   {ins:16,d}")
    print(f"\ntop {top} lines")
    for s, ins, f, ln, src in sorted(lines, reverse=True)[:top]:
        print(f"{s:8d} {100 * s / tot:5.1f}%  {ins:14,d}  {f}:{ln}  {src}")


if __name__ == "__main__":
    main()

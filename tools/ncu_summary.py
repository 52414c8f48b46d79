#!/usr/bin/env python3
"""Write a text summary of an ncu report (the numbers DESIGN.md / bench.py cite) — run here, no GPU needed.

usage: tools/ncu_summary.py REPORT.ncu-rep WARPS SAMPLES_PER_STREAM OUT.txt [--traffic-json OUT.json STREAMS SECONDS]
                            [--kernel REGEX] [--skip N]     (pick one launch of a multi-kernel report)
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep, warps, samples, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    sel = []
    if "--kernel" in sys.argv:
        sel += ["--kernel-name", "regex:" + sys.argv[sys.argv.index("--kernel") + 1]]
    if "--skip" in sys.argv:
        sel += ["--launch-skip", sys.argv[sys.argv.index("--skip") + 1]]
    sel += ["--launch-count", "1"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    got = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
    lines = [f"ncu summary of {rep}", f"kernel: {got.get('Kernel Name', ('?', ''))[0]}", ""]
    for m in METRICS:
        if m in got:
            lines.append(f"{m:75s} {got[m][0]:>18s} {got[m][1]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h = srows[1]
    ie, ss = h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = {n: h.index(n) for n in h if n.startswith("stall_") and "Not Issued" not in n}
    data = [r for r in srows[2:] if len(r) > ie and r[ie].isdigit()]
    tot = sum(int(r[ie]) for r in data)
    raw_inst = float(got["smsp__inst_executed.sum"][0].replace(",", "")) if "smsp__inst_executed.sum" in got else float(tot)
    lines += ["", f"executed warp-instructions {raw_inst:,.0f} = {raw_inst / (warps * samples):.1f} per sample step ({warps} warps x {samples} samples)"]
    agg = {}
    for r in data:
        for n, c in stall_cols.items():
            if r[c].isdigit():
                agg[n] = agg.get(n, 0) + int(r[c])
    tots = sum(agg.values()) or 1
    lines.append("warp stall samples: " + ", ".join(f"{k[6:]} {v / tots * 100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    ops = {}
    for r in data:
        t = r[1].strip().split()
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ie])
    lines.append("instruction mix: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if "--traffic-json" in sys.argv:
        i = sys.argv.index("--traffic-json")
        jout, streams, seconds = sys.argv[i + 1], int(sys.argv[i + 2]), float(sys.argv[i + 3])

        def num(name):
            v, u = got[name]
            f = float(v.replace(",", ""))
            return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

        import datetime
        import hashlib
        import os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        hh = hashlib.sha256()
        for f in ("same_kernels.cu", "same_fast.cuh", "same_lane.cuh", "same_transport.cuh", "same_params.h"):
            hh.update(open(os.path.join(root, "sameold_b200", "csrc", f), "rb").read())
        json.dump({"streams": streams, "seconds": seconds, "report": os.path.basename(rep),
                   "captured": datetime.date.today().isoformat(), "kernel_src_sha16": hh.hexdigest()[:16],
                   "kernel": got.get("Kernel Name", ("?", ""))[0].split("(")[0],
                   "dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
                   "dram_bytes_read": int(num("dram__bytes_read.sum")), "dram_bytes_write": int(num("dram__bytes_write.sum"))},
                  open(jout, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Measured A/B of DESIGN.md §5: fused receiver kernels (raw s16 read once, integer DC blocker inside the loop kernel)
against the split pipeline (time-parallel front-end kernel: s16 -> exact DC-blocked f32 lane-major tiles, then the
single-warp loop kernel fed from those tiles), plus the front-end kernel alone against the HBM roofline.

usage: tools/ab_split.py STREAMS SECONDS [STREAMS SECONDS ...]   -> one JSON line per workload
All times are CUDA-event times of the engine (same_engine_last_timing / same_engine_frontend_probe), device-resident
inputs, every step from a freshly reset receiver; best of `--reps` after one warm-up.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RATE = 22050


def main():
    import torch
    import sameold_b200 as sb
    from sameold_b200 import synth
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    reps = 3
    peak = 6541.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    for i in range(0, len(args), 2):
        ns, secs = int(args[i]), float(args[i + 1])
        n = int(secs * RATE)
        stride = (n + 7) // 8 * 8
        buf = torch.empty((ns, stride), dtype=torch.int16, device="cuda")
        plans = synth.plan_corpus(ns, RATE, secs)
        synth.generate_on_device(plans, buf.data_ptr(), stride, n, RATE)
        offsets = np.arange(ns, dtype=np.uint64) * np.uint64(stride)
        lengths = np.full(ns, n, np.uint32)
        rx = sb.SameReceiverBuilder.samedec(RATE).build_batch(ns)
        out = {"streams": ns, "seconds": secs, "samples": ns * n, "policy_kernel": rx.get_option("kernel_selected")}
        ref = None
        cases = [("fused_single_warp", 2, 0), ("fused_lookahead", 6, 0), ("fused_three_warp", 4, 0), ("fused_pipelined", 3, 0),
                 ("split_frontend_plus_tilefed", 5, 0)]
        if "--quick" in sys.argv:
            cases = cases[:2]
        for name, kernel, variant in cases:
            rx.set_option("kernel", kernel)
            best = None
            for r in range(reps + 1):
                rx.reset()
                rx.submit_device(buf.data_ptr(), ns * stride, offsets, lengths)
                rx.sync()
                ev, pay = rx.drain_raw(reuse=True)
                ms = rx.last_timing()[1]
                if r:
                    best = ms if best is None else min(best, ms)
            key = (int(ev.size), int((ev["kind"] == 18).sum()), int(ev["sample"].sum() % (1 << 61)))
            ref = ref or key
            assert key == ref, f"{name}: decodes differently {key} vs {ref}"
            out[name + "_ms"] = round(best, 3)
        fe = rx.frontend_probe(buf.data_ptr(), ns * stride, offsets, lengths, reps=5)
        out["frontend_alone_ms"] = round(fe, 3)
        out["frontend_gbs"] = round(6.0 * ns * n / (fe * 1e-3) / 1e9, 1)     # 2 B read + 4 B written per sample
        out["frontend_frac_of_hbm_peak"] = round(out["frontend_gbs"] / peak, 4)
        out["hbm_peak_gbs"] = peak
        out["events"], out["headers"] = ref[0], ref[1]
        print(json.dumps(out), flush=True)
        del rx, buf
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

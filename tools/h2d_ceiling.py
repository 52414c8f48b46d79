#!/usr/bin/env python3
"""Bare pinned host->device copy ceiling of a box: one host thread (or process) + one pinned buffer per GPU, all GPUs of
a subset copying at once, no kernel (same_h2d_probe of libsame_b200.so: CUDA-event timed cudaMemcpy2DAsync loop).

usage: tools/h2d_ceiling.py [--mb 2048] [--reps 8] [--bind] [--procs] [--rows R --width W --pitch P] SUBSET [SUBSET ...]
  SUBSET = comma-separated device ids.  Default pattern: one flat copy of --mb MiB per rep.  --rows/--width/--pitch
  (bytes): the strided pattern of same_engine_submit_s16_2d (R rows of W bytes, source pitch P; buffer = R * P bytes).
  --procs: one PROCESS per GPU (as under torchrun) instead of one thread per GPU in one process.
  --bind pins each worker to an even share of the host cores before it allocates its pinned buffer (first touch).
prints one JSON line per subset: per-device GB/s, aggregate GB/s (total bytes / max time).
"""
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def work(k, d, ndev, cfg, start, out):
    from sameold_b200 import _lib
    lib = _lib.load()
    if cfg["bind"]:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // ndev)
        os.sched_setaffinity(0, set(cores[k * per:(k + 1) * per]) or set(cores))
    rows, width, pitch = cfg["rows"], cfg["width"], cfg["pitch"]
    nbytes = rows * pitch
    h = lib.same_host_alloc(nbytes)
    C.memset(C.c_void_p(h), 1, nbytes)
    ms = C.c_float()
    start.wait()
    rc = lib.same_h2d_probe(d, C.c_void_p(h), pitch, width, rows, cfg["reps"], C.byref(ms))
    out[d] = (rc, ms.value)
    lib.same_host_free(C.c_void_p(h))


def main():
    args = sys.argv[1:]
    cfg = {"mb": 2048, "reps": 8, "bind": False, "procs": False, "rows": 0, "width": 0, "pitch": 0}
    subsets = []
    i = 0
    while i < len(args):
        a = args[i]
        if a in ("--mb", "--reps", "--rows", "--width", "--pitch"):
            cfg[a[2:]] = int(args[i + 1]); i += 2
        elif a in ("--bind", "--procs"):
            cfg[a[2:]] = True; i += 1
        else:
            subsets.append([int(x) for x in a.split(",")]); i += 1
    if not cfg["rows"]:
        cfg["rows"], cfg["width"], cfg["pitch"] = 1, cfg["mb"] << 20, cfg["mb"] << 20
    moved = cfg["rows"] * cfg["width"] * cfg["reps"]
    for devs in subsets:
        if cfg["procs"]:
            ctx = mp.get_context("spawn")
            mgr = ctx.Manager()
            res = mgr.dict()
            start = ctx.Barrier(len(devs))
            ws = [ctx.Process(target=work, args=(k, d, len(devs), cfg, start, res)) for k, d in enumerate(devs)]
        else:
            res = {}
            start = threading.Barrier(len(devs))
            ws = [threading.Thread(target=work, args=(k, d, len(devs), cfg, start, res)) for k, d in enumerate(devs)]
        [w.start() for w in ws]
        [w.join() for w in ws]
        res = dict(res)
        worst = max(v[1] for v in res.values())
        print(json.dumps({"devices": devs, "workers": "processes" if cfg["procs"] else "threads", "bind": cfg["bind"],
                          "rows": cfg["rows"], "width_bytes": cfg["width"], "pitch_bytes": cfg["pitch"],
                          "buffer_mb": cfg["rows"] * cfg["pitch"] >> 20, "reps": cfg["reps"],
                          "per_device_gbs": {d: round(moved / (v[1] * 1e-3) / 1e9, 2) for d, v in sorted(res.items())},
                          "aggregate_gbs": round(len(devs) * moved / (worst * 1e-3) / 1e9, 2),
                          "rc": [v[0] for v in res.values()]}), flush=True)


if __name__ == "__main__":
    main()

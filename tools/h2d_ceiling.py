#!/usr/bin/env python3
"""Bare pinned host->device copy ceiling of a box: one host thread + one pinned buffer per GPU, all GPUs of a subset
copying at once, no kernel (same_h2d_probe of libsame_b200.so: CUDA-event timed cudaMemcpy2DAsync loop).

usage: tools/h2d_ceiling.py [--mb 2048] [--reps 8] [--bind] SUBSET [SUBSET ...]     SUBSET = comma-separated device ids
prints one JSON line per subset: per-device GB/s, aggregate GB/s (total bytes / max time).
--bind pins each thread to an even share of the host cores before it allocates its pinned buffer (first touch).
"""
import ctypes as C
import json
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from sameold_b200 import _lib
    lib = _lib.load()
    args = sys.argv[1:]
    mb, reps, bind = 2048, 8, False
    subsets = []
    i = 0
    while i < len(args):
        if args[i] == "--mb":
            mb = int(args[i + 1]); i += 2
        elif args[i] == "--reps":
            reps = int(args[i + 1]); i += 2
        elif args[i] == "--bind":
            bind = True; i += 1
        else:
            subsets.append([int(x) for x in args[i].split(",")]); i += 1
    nbytes = mb << 20
    cores = sorted(os.sched_getaffinity(0))
    for devs in subsets:
        res = {}
        start = threading.Barrier(len(devs))

        def work(k, d):
            if bind:
                per = max(1, len(cores) // len(devs))
                os.sched_setaffinity(0, set(cores[k * per:(k + 1) * per]) or set(cores))
            h = lib.same_host_alloc(nbytes)
            C.memset(C.c_void_p(h), 1, nbytes)
            ms = C.c_float()
            start.wait()
            rc = lib.same_h2d_probe(d, C.c_void_p(h), nbytes, nbytes, 1, reps, C.byref(ms))
            res[d] = (rc, ms.value)
            lib.same_host_free(C.c_void_p(h))

        th = [threading.Thread(target=work, args=(k, d)) for k, d in enumerate(devs)]
        [t.start() for t in th]
        [t.join() for t in th]
        worst = max(v[1] for v in res.values())
        print(json.dumps({"devices": devs, "bind": bind, "mb_per_copy": mb, "reps": reps,
                          "per_device_gbs": {d: round(nbytes * reps / (v[1] * 1e-3) / 1e9, 2) for d, v in sorted(res.items())},
                          "aggregate_gbs": round(len(devs) * nbytes * reps / (worst * 1e-3) / 1e9, 2),
                          "rc": [v[0] for v in res.values()]}), flush=True)


if __name__ == "__main__":
    main()

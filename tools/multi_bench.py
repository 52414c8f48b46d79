#!/usr/bin/env python3
"""The in-library multi-device entry (same_multi_*: one host thread + one engine per GPU inside ONE process) on
BASELINE config 4's shape: STREAMS (65536) synthetic streams sharded contiguously over the visible GPUs, fed from one
pinned host matrix in time-chunks through same_multi_submit_s16_2d, events gathered with global stream ids.

usage: tools/multi_bench.py [--streams 65536] [--seconds 20] [--chunk-seconds 5] [--devices 0,1,...]
prints one JSON line: host wall-clock rate (end to end from host memory) and the per-device CUDA-event times.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
RATE = 22050


def main():
    import torch
    import sameold_b200 as sb
    from sameold_b200 import synth, _lib
    a = sys.argv[1:]
    opt = {"--streams": "65536", "--seconds": "20", "--chunk-seconds": "5", "--devices": ""}
    for i in range(0, len(a), 2):
        opt[a[i]] = a[i + 1]
    ns, secs, cs = int(opt["--streams"]), float(opt["--seconds"]), float(opt["--chunk-seconds"])
    devices = [int(x) for x in opt["--devices"].split(",")] if opt["--devices"] else list(range(torch.cuda.device_count()))
    lib = _lib.load()
    n = int(secs * RATE)
    cn = int(cs * RATE) // 8 * 8
    nch = (n + cn - 1) // cn
    b = sb.SameReceiverBuilder.samedec(RATE)
    rx = b.build_multi(ns, devices)
    shards = rx.shards()
    # corpus: each device generates its shard's chunk, which is then staged in the one pinned host matrix (untimed)
    hptr = lib.same_host_alloc(ns * cn * 2)
    host = torch.from_numpy(np.ctypeslib.as_array(C.cast(hptr, C.POINTER(C.c_int16)), shape=(ns, cn)))
    gens = []
    for dev, first, count in shards:
        with torch.cuda.device(dev):
            plans = synth.plan_corpus(count, RATE, secs, first_stream=first)
            gens.append((synth.DeviceCorpus(plans, RATE, device=dev), torch.empty((count, cn), dtype=torch.int16, device=f"cuda:{dev}")))
    nsl = 4
    wall, headers, events = 0.0, 0, 0
    for c in range(nch):
        w = min(cn, n - c * cn)
        for (dev, first, count), (corpus, buf) in zip(shards, gens):
            corpus.generate(buf.data_ptr(), cn, w, first_sample=c * cn)
            host[first:first + count].copy_(buf)
        for dev in devices:
            torch.cuda.synchronize(dev)
        cuts = [int(round(i * w / nsl)) for i in range(nsl + 1)]
        t0 = time.perf_counter()
        for i in range(nsl):
            rx.submit_2d(hptr, cn, cuts[i], cuts[i + 1] - cuts[i])
        rx.sync()
        ev, pay = rx.drain_raw()
        wall += time.perf_counter() - t0
        headers += int((ev["kind"] == 18).sum())
        events += int(ev.size)
        assert ev.size == 0 or (np.all(np.diff(ev["stream"].astype(np.int64)) >= 0) and int(ev["stream"].max()) < ns)
    audio = ns * secs
    print(json.dumps({"api": "same_multi_submit_s16_2d + same_multi_sync + same_multi_drain_events (one process, one host thread "
                             "+ engine per device)", "devices": devices, "streams": ns, "seconds": secs, "chunk_seconds": cn / RATE,
                      "wall_s": round(wall, 4), "audio_s_per_s_e2e_host_wall": round(audio / wall, 1),
                      "h2d_gbs": round(ns * n * 2 / wall / 1e9, 2), "headers": headers, "events": events,
                      "shards": shards}), flush=True)
    assert secs < 60 or headers >= int(0.95 * ns)
    lib.same_host_free(C.c_void_p(hptr))


if __name__ == "__main__":
    main()

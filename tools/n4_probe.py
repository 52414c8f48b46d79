#!/usr/bin/env python3
"""torchrun rank script: what does a rank's environment do to its host->device rate?  Each rank allocates the e2e
pinned pool (4096 rows x 2646016 B), then all ranks run same_h2d_probe at once (bench.py's strided pattern).
  --backend nccl|gloo|none   process group used for the barrier (none: file-less, ranks just start together)
  --torch-cuda 0|1           whether torch creates its CUDA context / caching allocator on the device first
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    a = sys.argv[1:]
    opt = {"--backend": "nccl", "--torch-cuda": "1", "--big-device-buffer": "0"}
    for i in range(0, len(a), 2):
        opt[a[i]] = a[i + 1]
    rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    from sameold_b200 import _lib
    lib = _lib.load()
    import torch
    import torch.distributed as dist
    big = None
    if opt["--torch-cuda"] == "1":
        torch.cuda.set_device(lr)
        torch.zeros(1, device="cuda")
        if opt["--big-device-buffer"] == "1":
            big = torch.empty((4096, 1323008), dtype=torch.int16, device="cuda")
    if opt["--backend"] == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    elif opt["--backend"] == "gloo":
        dist.init_process_group("gloo")
    rows, width, pitch = 4096, 110250, 2646016
    h = lib.same_host_alloc(rows * pitch)
    C.memset(C.c_void_p(h), 1, rows * pitch)
    if big is not None:   # fill the pinned pool the way bench.py does: a device->host copy
        import numpy as np
        host = torch.from_numpy(np.ctypeslib.as_array(C.cast(h, C.POINTER(C.c_int16)), shape=(rows, pitch // 2)))
        host.copy_(big)
        torch.cuda.synchronize()
    out = []
    for rep in range(2):
        if opt["--backend"] != "none":
            dist.barrier()
        ms = C.c_float()
        rc = lib.same_h2d_probe(lr, C.c_void_p(h), pitch, width, rows, 24, C.byref(ms))
        out.append(round(rows * width * 24 / (ms.value * 1e-3) / 1e9, 2))
    print(json.dumps({"rank": rank, "world": world, "opt": opt, "gbs": out, "omp": os.environ.get("OMP_NUM_THREADS"),
                      "cvd": os.environ.get("CUDA_VISIBLE_DEVICES")}), flush=True)
    if opt["--backend"] != "none":
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/bin/bash
# round-2 call B: parity of the refactored kernels + split pipeline, A/B table, ncu of the front-end and fast kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k "each_fast_kernel or fast_and_generic or config3_full or config4 or golden or tiny or large_batches" --durations=8 ) > gpurun_out/r2b_tests.txt 2>&1
tail -4 gpurun_out/r2b_tests.txt
( time timeout 900 python tools/ab_split.py 4096 60 16384 20 65536 20 ) > gpurun_out/r2b_ab.jsonl 2> gpurun_out/r2b_ab.err
cat gpurun_out/r2b_ab.jsonl; tail -3 gpurun_out/r2b_ab.err
# ncu: front-end kernel + fused fast kernel (both variants) at 65536 x 5 s
cat > /tmp/ncu_job.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import sameold_b200 as sb
from sameold_b200 import synth
ns, secs = 65536, 5.0
n = int(secs * 22050); stride = (n + 7) // 8 * 8
buf = torch.empty((ns, stride), dtype=torch.int16, device="cuda")
synth.generate_on_device(synth.plan_corpus(ns, 22050, secs), buf.data_ptr(), stride, n)
off = np.arange(ns, dtype=np.uint64) * np.uint64(stride); ln = np.full(ns, n, np.uint32)
rx = sb.SameReceiverBuilder.samedec(22050).build_batch(ns)
for kernel, variant in ((2, 0), (2, 1), (5, 0)):
    rx.set_option("kernel", kernel); rx.set_option("fast_variant", variant)
    rx.reset(); rx.submit_device(buf.data_ptr(), ns * stride, off, ln); rx.sync(); rx.drain_raw()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'same_rx_fast|same_frontend' -c 5 -o gpurun_out/r2b_fast_65536x5 python /tmp/ncu_job.py > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log

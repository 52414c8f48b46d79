#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_engine_gpu.py -x -q -k "each_fast_kernel or fast_and_generic or submit_2d_odd" ) > gpurun_out/r2e_tests.txt 2>&1
tail -4 gpurun_out/r2e_tests.txt
( time timeout 900 python tools/ab_split.py --quick 8192 20 49152 20 65536 20 131072 10 ) > gpurun_out/r2e_sweep.jsonl 2> gpurun_out/r2e_sweep.err
cat gpurun_out/r2e_sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['streams'], d['seconds'], 'fast', d['fused_single_warp_ms'], 'lookahead', d['fused_lookahead_ms'])"
tail -3 gpurun_out/r2e_sweep.err
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
rx.set_option("kernel", 6)
rx.reset(); rx.submit_device(buf.data_ptr(), ns * stride, off, ln); rx.sync(); rx.drain_raw()
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'same_rx_la' -c 1 -o gpurun_out/r2e_la_65536x5 python /tmp/ncu_job.py > gpurun_out/r2e_ncu.log 2>&1
tail -2 gpurun_out/r2e_ncu.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ncu_job.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import sameold_b200 as sb
from sameold_b200 import synth
rate = 22050; n = 300 * rate
plan = synth.plan_long_stream(0.2, rate)
buf = torch.empty(((n + 7) // 8 * 8,), dtype=torch.int16, device="cuda")
synth.DeviceCorpus([plan], rate).generate(buf.data_ptr(), buf.numel(), n)
rx = sb.SameReceiverBuilder.samedec(rate).build_batch(1)
rx.submit_device(buf.data_ptr(), n, np.zeros(1, np.uint64), np.array([n], np.uint32)); rx.sync(); print(len(rx.drain()), rx.last_timing())
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'same_long_seq' -c 1 -o gpurun_out/r2l_seq python /tmp/ncu_job.py > gpurun_out/r2l_ncu.log 2>&1; tail -3 gpurun_out/r2l_ncu.log

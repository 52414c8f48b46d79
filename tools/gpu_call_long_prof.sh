#!/bin/bash
# profiles of the long-stream path: launch list of a 1 h single-stream run, ncu --set full of the sequential kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2p_long_launches.csv python bench.py --config 5 --hours 1 --no-cpu > gpurun_out/r2p_long_launches.log 2>&1; tail -1 gpurun_out/r2p_long_launches.log | cut -c1-200
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
rx.submit_device(buf.data_ptr(), n, np.zeros(1, np.uint64), np.array([n], np.uint32)); rx.sync(); print(len(rx.drain()))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'same_long' -c 5 -o gpurun_out/r2p_long python /tmp/ncu_job.py > gpurun_out/r2p_ncu.log 2>&1; tail -2 gpurun_out/r2p_ncu.log

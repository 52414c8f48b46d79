#!/bin/bash
# round-2 final validation: GPU tests, smoke, bench (config 3 + config-4 leg, config 5), ncu launch list and the config-3
# traffic capture (profiles/rx_kernel_traffic.json via tools/ncu_summary.py --traffic-json)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 ) > gpurun_out/r2z_tests.txt 2>&1
tail -12 gpurun_out/r2z_tests.txt
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2z_smoke.txt 2>&1; grep -i smoke gpurun_out/r2z_smoke.txt
( time timeout 600 python bench.py ) > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; tail -c 300 gpurun_out/r2z_bench.json; tail -2 gpurun_out/r2z_bench.err
( time timeout 600 python bench.py --config 5 --hours 24 ) > gpurun_out/r2z_config5.json 2> gpurun_out/r2z_config5.err; tail -c 600 gpurun_out/r2z_config5.json; tail -2 gpurun_out/r2z_config5.err
if [ "$1" = "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2z_launches.log 2>&1; tail -1 gpurun_out/r2z_launches.log | cut -c1-200
cat > /tmp/ncu_job.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import sameold_b200 as sb
from sameold_b200 import synth
ns, secs = 4096, 60.0
n = int(secs * 22050); stride = (n + 7) // 8 * 8
buf = torch.empty((ns, stride), dtype=torch.int16, device="cuda")
synth.generate_on_device(synth.plan_corpus(ns, 22050, secs), buf.data_ptr(), stride, n)
off = np.arange(ns, dtype=np.uint64) * np.uint64(stride); ln = np.full(ns, n, np.uint32)
rx = sb.SameReceiverBuilder.samedec(22050).build_batch(ns)
rx.reset(); rx.submit_device(buf.data_ptr(), ns * stride, off, ln); rx.sync(); rx.drain_raw()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'same_rx_pipe' -c 1 -o gpurun_out/r2z_pipe_config3 python /tmp/ncu_job.py > gpurun_out/r2z_ncu1.log 2>&1; tail -1 gpurun_out/r2z_ncu1.log
fi

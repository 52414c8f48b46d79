#!/bin/bash
# what halves the per-GPU host->device rate at 4 ranks?  (in-bench ceiling 114 GB/s against 222 GB/s for flat 2 GB copies)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S="0,1,2,3"
{
python tools/h2d_ceiling.py --mb 2048 $S
python tools/h2d_ceiling.py --mb 10240 --reps 3 $S
python tools/h2d_ceiling.py --rows 4096 --width 110250 --pitch 2646016 --reps 24 $S
python tools/h2d_ceiling.py --rows 4096 --width 110250 --pitch 110592 --reps 24 $S
python tools/h2d_ceiling.py --rows 4096 --width 2646016 --pitch 2646016 --reps 2 $S
python tools/h2d_ceiling.py --procs --mb 2048 $S
python tools/h2d_ceiling.py --procs --rows 4096 --width 110250 --pitch 2646016 --reps 24 $S
python tools/h2d_ceiling.py --procs --bind --rows 4096 --width 110250 --pitch 2646016 --reps 24 $S
python tools/h2d_ceiling.py --rows 4096 --width 110250 --pitch 2646016 --reps 24 0,1
python tools/h2d_ceiling.py --rows 4096 --width 110250 --pitch 2646016 --reps 24 0
} > gpurun_out/r2n_h2d.jsonl 2> gpurun_out/r2n_h2d.err
python - <<'PY'
import json
for l in open("gpurun_out/r2n_h2d.jsonl"):
    d = json.loads(l); print(d["workers"], d["bind"], d["rows"], d["width_bytes"], d["pitch_bytes"], d["buffer_mb"], "MB", d["devices"], "->", d["aggregate_gbs"], d["per_device_gbs"])
PY
tail -3 gpurun_out/r2n_h2d.err

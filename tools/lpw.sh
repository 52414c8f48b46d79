cd "$(dirname "$0")/.."
# lanes-per-warp sweep of the pipelined kernel on config 3 (4096 streams on 148 SMs: 32 lanes = 128 blocks, 28 = 147)
for l in 32 28 30 32 28; do
  out=$(timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-config4 --lanes-per-warp $l 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['roofline']['kernel_ms_per_launch'], d['ms_per_step'])")
  echo "lanes=$l => $out"
done

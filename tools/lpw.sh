cd /root/repo
for k in 3 4; do for l in 32 16 8; do
  out=$(timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-config4 --kernel $k --lanes-per-warp $l 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['roofline']['kernel_ms_per_launch'], d['ms_per_step'])")
  echo "kernel=$k lanes=$l => $out"
done; done

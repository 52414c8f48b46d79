#!/bin/bash
# A/B of the long-stream ring hand-off: release/acquire (default build) vs volatile (-DLS_RELEASE_ACQUIRE=0 variant),
# then the long-stream parity tests on the default build.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/sameold_b200/_build/libsame_b200_volatile_ring.so
for rep in 1 2; do
  for lib in "" "$V"; do
    out=$(SAME_B200_LIB=$lib timeout 200 python bench.py --config 5 --hours 2 --no-cpu 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d.get('ms_per_step'))")
    echo "lib=${lib:-default(release/acquire)} => $out" | tee -a gpurun_out/ring_ab.txt
  done
done
timeout 400 python -m pytest tests -x -q -m gpu -k "long or 24h" 2>&1 | tail -3 | tee -a gpurun_out/ring_ab.txt
timeout 300 python bench.py --config 5 2>gpurun_out/ring_c5.err | grep '^{' > gpurun_out/bench_config5_24h.json; cut -c1-260 gpurun_out/bench_config5_24h.json

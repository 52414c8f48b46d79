#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/r2g_variants.txt
for lib in libv1.so libv6.so libv7.so libv8.so; do
  out=$(SAME_B200_LIB=$PWD/sameold_b200/_build/$lib timeout 600 python tools/ab_split.py --quick 49152 20 65536 20 131072 10 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['streams'], 'fast', d['fused_single_warp_ms'], 'la', d['fused_lookahead_ms'], end=' | ')")
  echo "$lib: $out" | tee -a gpurun_out/r2g_variants.txt
done

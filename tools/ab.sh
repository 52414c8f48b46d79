#!/bin/bash
# A/B timing of in-tree builds on one box: tools/ab.sh "<lib or ->:<kernel option (1-4) or ->:<bench args>" ...
# prints kernel ms per launch for each case, twice (ABAB) to expose drift.
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for spec in "$@"; do
    IFS=: read -r lib kern args <<<"$spec"
    env=()
    [ "$lib" != "-" ] && env+=("SAME_B200_LIB=$PWD/sameold_b200/_build/$lib")
    kopt=""; [ "$kern" != "-" ] && kopt="--kernel $kern"
    out=$(env "${env[@]}" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-config4 $kopt $args 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['roofline']['kernel_ms_per_launch'], d['ms_per_step'])")
    echo "$spec => $out"
  done
done

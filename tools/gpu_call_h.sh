#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=12 ) > gpurun_out/r2h_tests.txt 2>&1
tail -30 gpurun_out/r2h_tests.txt
( time timeout 600 python bench.py ) > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -c 1500 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err

#!/bin/bash
# round-2 call A: new parity tests, bench with config-4 leg, config-5 24 h run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; free -g >> gpurun_out/r2a_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_config5_24h.py --durations=15 ) > gpurun_out/r2a_tests.txt 2>&1
( time timeout 600 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
( time timeout 600 python bench.py --config 5 --hours 24 ) > gpurun_out/r2a_config5.json 2> gpurun_out/r2a_config5.err
( time SAME_TEST_HOURS=2 timeout 600 python -m pytest tests/test_zz_config5_24h.py -q -s ) > gpurun_out/r2a_test5_2h.txt 2>&1
tail -5 gpurun_out/r2a_tests.txt; tail -c 600 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err; tail -c 400 gpurun_out/r2a_config5.json; tail -5 gpurun_out/r2a_test5_2h.txt

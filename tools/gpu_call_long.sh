#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -k "long_stream_path or long_single_stream" ) > gpurun_out/r2l_tests.txt 2>&1
tail -25 gpurun_out/r2l_tests.txt
( time SAME_TEST_HOURS=2 timeout 600 python -m pytest tests/test_zz_config5_24h.py -x -q -s ) > gpurun_out/r2l_2h.txt 2>&1
tail -8 gpurun_out/r2l_2h.txt

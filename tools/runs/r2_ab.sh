#!/bin/bash
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_paths.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/r2ab_tests.log
{ python tools/quick_bench.py 2048 128 0 5; python tools/quick_bench.py 1024 512 1 3; python tools/pass_times.py 2048 32; python tools/pass_times.py 1024 64 1; } > gpurun_out/r2ab_quick.log 2>&1

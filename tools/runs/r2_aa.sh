#!/bin/bash
mkdir -p gpurun_out
{ python tools/quick_bench.py 512 1024 1 10; python tools/quick_bench.py 1024 512 0 10; python tools/quick_bench.py 512 256 0 10; } > gpurun_out/r2aa_quick.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_paths.py -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2aa_tests.log

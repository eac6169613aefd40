#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12) > gpurun_out/r2ah_tests.log
{ python tools/quick_bench.py 128 8192 1 10; python tools/quick_bench.py 256 2048 1 10; python tools/quick_bench.py 2048 32 1 2; } > gpurun_out/r2ah_quick.log 2>&1

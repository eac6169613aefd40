#!/bin/bash
# transposed-intermediate FFT-1024 kernels: full GPU tests, then bench c3 / c2-like shapes that use FFT 1024
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2i_tests.log
python bench.py --config c3 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline --no-e2e > gpurun_out/r2i_bench_c3.json 2> gpurun_out/r2i_bench_c3.err
python tools/quick_bench.py 512 1024 1 10 > gpurun_out/r2i_quick.log 2>&1

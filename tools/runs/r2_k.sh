#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2k_tests.log
python bench.py --config c3 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline --no-e2e > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err
python tools/quick_bench.py 512 1024 1 10 > gpurun_out/r2k_quick.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k32t -c 12 --csv --log-file gpurun_out/r2k_launches.csv python tools/prof_case.py 1024 27 0 1 > gpurun_out/r2k_ncu.log 2>&1

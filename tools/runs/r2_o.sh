#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2o_tests.log
{ python tools/quick_bench.py 1024 512 0 10; python tools/quick_bench.py 512 1024 1 10; python tools/pass_times.py 1024 108; } > gpurun_out/r2o_quick.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/r2o_dram_c3.csv python tools/prof_case.py 1024 108 0 1 > gpurun_out/r2o_dram_c3.log 2>&1

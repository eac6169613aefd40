#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2r_tests.log
{
echo "== product"; python tools/quick_bench.py 1024 512 0 10; python tools/quick_bench.py 512 1024 1 10; python tools/pass_times.py 1024 108
export ASM_B200_LIB=$D/libasm_b200_tune.so
echo "== lanes/chunk sweep"
for l in 2 3 4; do for mb in 72 96 120 144 216 288; do ASM_B200_LANES=$l ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 1024 512 0 10; done; done
} > gpurun_out/r2r_sweep.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/r2r_dram_c3.csv python tools/prof_case.py 1024 108 0 1 > gpurun_out/r2r_dram_c3.log 2>&1

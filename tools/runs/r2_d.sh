#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
export ASM_B200_LIB=$D/libasm_b200_tune.so
echo "== 1024 default (6-warp rows 2/SM)"; python tools/quick_bench.py 1024 512
echo "== 2048 CC=4, 2 CTAs/SM"; python tools/quick_bench.py 2048 128; python tools/pass_times.py 2048 32; python tools/quick_bench.py 1024 128 1
echo "== 2048 chunk/lanes"; for mb in 144 288 432; do for l in 2 3 4; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 2048 128; done; done
export ASM_B200_LIB=$D/libasm_b200_tune_cc8.so
echo "== 2048 CC=8 1 CTA/SM"; python tools/quick_bench.py 2048 128; python tools/pass_times.py 2048 32; python tools/quick_bench.py 1024 128 1
export ASM_B200_LIB=$D/libasm_b200_tune.so
echo "== 256 chunked sweeps"
for mb in 48 96 144 216; do for l in 3 4 6; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 256 4096; done; done
echo "== 128 pad chunked sweeps"
for mb in 48 96 144; do for l in 3 4; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 128 8192 1; done; done
echo "== 512 (FFT 512) sweeps: chunked vs resident"
for mb in 48 96 144; do ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 512 1024; ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 256 2048 1; done
ASM_B200_RESIDENT=1 python tools/quick_bench.py 512 1024; ASM_B200_RESIDENT=1 python tools/quick_bench.py 256 2048 1
} > gpurun_out/r2d_sweep.log 2>&1

#!/bin/bash
# round-2 batch E: FFT-2048 warp-pair kernels (k64.cuh): parity suite + A/B against the generic kernels
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2e_tests.log
D=$PWD/style_transfer_based_holographic_imaging_b200
{
export ASM_B200_LIB=$D/libasm_b200_tune.so
for k in 1 0; do
  echo "== K64=$k"
  ASM_B200_K64=$k python tools/quick_bench.py 2048 128; ASM_B200_K64=$k python tools/pass_times.py 2048 32
  ASM_B200_K64=$k python tools/quick_bench.py 1024 128 1; ASM_B200_K64=$k python tools/pass_times.py 1024 32 1
done
echo "== K64 chunk/lanes"
for mb in 144 216 288 432; do for l in 2 3 4; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 2048 128; done; done
for mb in 144 216 288; do for l in 2 3; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 1024 128 1; done; done
} > gpurun_out/r2e_sweep.log 2>&1

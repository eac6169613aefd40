#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
export ASM_B200_LIB=$D/libasm_b200_tune.so
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_paths.py -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2y_tests.log
{
for k in 1 0; do
  echo "== K64T=$k"
  ASM_B200_K64T=$k python tools/quick_bench.py 2048 128 0 5
  ASM_B200_K64T=$k python tools/quick_bench.py 1024 512 1 3
  ASM_B200_K64T=$k python tools/pass_times.py 2048 32
  ASM_B200_K64T=$k python tools/pass_times.py 1024 64 1
done
} > gpurun_out/r2y_k64t.log 2>&1

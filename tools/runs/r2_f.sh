#!/bin/bash
# round-2 batch F: launch-sequence (CUDA graph) cache A/B on the launch-bound small transforms + full test suite
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2f_tests.log
D=$PWD/style_transfer_based_holographic_imaging_b200
{
export ASM_B200_LIB=$D/libasm_b200_tune.so
for gph in 1 0; do
  echo "== GRAPHS=$gph"
  ASM_B200_GRAPHS=$gph python tools/quick_bench.py 256 4096 0 10
  ASM_B200_GRAPHS=$gph python tools/quick_bench.py 256 4096 1 10
  ASM_B200_GRAPHS=$gph python tools/quick_bench.py 128 8192 0 10
  ASM_B200_GRAPHS=$gph python tools/quick_bench.py 128 8192 1 10
  ASM_B200_GRAPHS=$gph python tools/quick_bench.py 512 1024 0 10
  ASM_B200_GRAPHS=$gph python tools/quick_bench.py 64 8192 1 10
  ASM_B200_GRAPHS=$gph python tools/split_bench.py 256 4096
done
echo "== GRAPHS=1 with larger n (1024, 2048)"
ASM_B200_GRAPH_MAX_N=11 python tools/quick_bench.py 1024 512 0 10
ASM_B200_GRAPH_MAX_N=11 python tools/quick_bench.py 2048 128 0 10
echo "== GRAPHS=1 chunk / lanes sweeps at 256"
for mb in 24 48 72; do for l in 3 4 6; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 256 4096 0 10; done; done
for mb in 24 48 72; do for l in 3 4 6; do ASM_B200_CHUNK_MB=$mb ASM_B200_LANES=$l python tools/quick_bench.py 128 8192 1 10; done; done
} > gpurun_out/r2f_sweep.log 2>&1
{
D=$PWD/style_transfer_based_holographic_imaging_b200
echo "== FFT 4096: CC=4 (1024 threads) vs CC=2 (512 threads, 2 CTAs/SM)"
ASM_B200_LIB=$D/libasm_b200_tune.so python tools/quick_bench.py 4096 32; ASM_B200_LIB=$D/libasm_b200_tune.so python tools/quick_bench.py 2048 32 1
ASM_B200_LIB=$D/libasm_b200_tune_cc12.so python tools/quick_bench.py 4096 32; ASM_B200_LIB=$D/libasm_b200_tune_cc12.so python tools/quick_bench.py 2048 32 1
ASM_B200_LIB=$D/libasm_b200_tune_cc12.so python tools/pass_times.py 4096 8
} > gpurun_out/r2f_4096.log 2>&1

#!/bin/bash
# round-2 batch B: full GPU test-suite on the lean build + A/B sweeps with the tuning build
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2b_tests.log
export ASM_B200_LIB=$PWD/style_transfer_based_holographic_imaging_b200/libasm_b200_tune.so
{
echo "== default (CC=8 BULK=3)"; python tools/quick_bench.py 1024 512; python tools/pass_times.py 1024 108
echo "== CC=4"; ASM_B200_COLS_CC=4 python tools/quick_bench.py 1024 512; ASM_B200_COLS_CC=4 python tools/pass_times.py 1024 108
echo "== BULK=7 (register stores)"; ASM_B200_BULK=7 python tools/quick_bench.py 1024 512; ASM_B200_BULK=7 python tools/pass_times.py 1024 108
echo "== BULK=7 CC=4"; ASM_B200_BULK=7 ASM_B200_COLS_CC=4 python tools/quick_bench.py 1024 512
echo "== BULK=0"; ASM_B200_BULK=0 python tools/quick_bench.py 1024 512; ASM_B200_BULK=0 python tools/pass_times.py 1024 108
for l in 4 6 8; do for mb in 96 144 216; do
  echo "== CC=4 LANES=$l CHUNK_MB=$mb"; ASM_B200_COLS_CC=4 ASM_B200_LANES=$l ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 1024 512
done; done
for l in 4 6; do for mb in 96 144; do
  echo "== CC=8 LANES=$l CHUNK_MB=$mb"; ASM_B200_LANES=$l ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 1024 512
done; done
echo "== padded 512 (FFT 1024)"; python tools/quick_bench.py 512 1024 1; ASM_B200_COLS_CC=4 python tools/quick_bench.py 512 1024 1
} > gpurun_out/r2b_sweep.log 2>&1

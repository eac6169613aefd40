#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2p_tests.log
{
echo "== product (2 row CTAs/SM, PF 1)"; python tools/quick_bench.py 1024 512 0 10; python tools/quick_bench.py 512 1024 1 10; python tools/pass_times.py 1024 108
for v in v20 v22 v31; do
  echo "== variant $v"; export ASM_B200_LIB=$D/libasm_b200_$v.so
  python tools/quick_bench.py 1024 512 0 10; python tools/pass_times.py 1024 108
done
export ASM_B200_LIB=$D/libasm_b200_v31.so
echo "== v31 lanes/chunk sweep"
for l in 2 3 4; do for mb in 144 216 324; do ASM_B200_LANES=$l ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 1024 512 0 10; done; done
} > gpurun_out/r2p_variants.log 2>&1

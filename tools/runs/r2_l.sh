#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
for v in A B C; do
  echo "== variant $v"; export ASM_B200_LIB=$D/libasm_b200_tune$v.so
  python tools/quick_bench.py 1024 512 0 10
  python tools/pass_times.py 1024 108
done
export ASM_B200_LIB=$D/libasm_b200_tuneA.so
echo "== A: old path"; ASM_B200_K32T=0 python tools/quick_bench.py 1024 512 0 10
echo "== A: register-landing fwd rows (BULK=2), inv (BULK=1), both (0)"
ASM_B200_BULK=2 python tools/quick_bench.py 1024 512 0 10
ASM_B200_BULK=1 python tools/quick_bench.py 1024 512 0 10
ASM_B200_BULK=0 python tools/quick_bench.py 1024 512 0 10
echo "== A: lanes / chunk sweeps"
for l in 2 3 4; do for mb in 108 216 324; do ASM_B200_LANES=$l ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 1024 512 0 10; done; done
} > gpurun_out/r2l_variants.log 2>&1

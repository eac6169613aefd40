#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
for v in 0 1 2 4; do
  echo "== L2 prefetch distance $v"; export ASM_B200_LIB=$D/libasm_b200_pf$v.so
  python tools/pass_times.py 1024 108
  python tools/quick_bench.py 1024 512 0 10
  python tools/quick_bench.py 512 1024 1 10
done
} > gpurun_out/r2n_pf.log 2>&1

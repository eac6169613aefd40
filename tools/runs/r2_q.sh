#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
for v in 0 5 6 7; do
  echo "== dbg $v (0 base, 5 lines copy only, 6 fwd rows copy only, 7 inv rows copy only)"; ASM_B200_LIB=$D/libasm_b200_dbg$v.so python tools/pass_times.py 1024 108
  ASM_B200_LIB=$D/libasm_b200_dbg$v.so python tools/quick_bench.py 1024 512 0 10
done
} > gpurun_out/r2q_dbg.log 2>&1

#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
echo "== base"; ASM_B200_LIB=$D/libasm_b200_tuneA.so python tools/pass_times.py 1024 108
for v in 1 2 3 4; do
  echo "== dbg $v (1 no tile store, 2 no forward row math, 3 no tile load, 4 no output store)"; ASM_B200_LIB=$D/libasm_b200_dbg$v.so python tools/pass_times.py 1024 108
done
} > gpurun_out/r2m_dbg.log 2>&1

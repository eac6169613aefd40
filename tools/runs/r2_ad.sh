#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
for v in v0 vA vB vC; do
  echo "== $v (v0 default; vA 3 row CTAs per SM; vB no L2 prefetch; vC prefetch 2 groups ahead)"; export ASM_B200_LIB=$D/libasm_b200_$v.so
  python tools/quick_bench.py 1024 512 0 10; python tools/quick_bench.py 512 1024 1 10
done
export ASM_B200_LIB=$D/libasm_b200_v0.so
echo "== v0: lanes x chunk fine sweep"
for cfg in "2 96" "2 112" "2 128" "2 160" "3 144" "3 168"; do set -- $cfg; ASM_B200_LANES=$1 ASM_B200_CHUNK_MB=$2 python tools/quick_bench.py 1024 512 0 10; done
} > gpurun_out/r2ad_variants.log 2>&1

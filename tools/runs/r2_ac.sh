#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
export ASM_B200_LIB=$D/libasm_b200_tune.so
{
echo "== 2048 unpadded (c4), 32 MB per sample"
for cfg in "3 216" "3 96" "2 128" "2 64" "1 64" "4 128" "3 144"; do set -- $cfg; ASM_B200_LANES=$1 ASM_B200_CHUNK_MB=$2 python tools/quick_bench.py 2048 128 0 5; done
echo "== 1024 padded (c3pad), 16 MB per sample"
for cfg in "3 216" "3 96" "2 128" "2 64" "3 144" "4 128"; do set -- $cfg; ASM_B200_LANES=$1 ASM_B200_CHUNK_MB=$2 python tools/quick_bench.py 1024 256 1 3; done
echo "== 2048 padded (c4pad), 64 MB per sample"
for cfg in "3 216" "2 128" "1 64" "2 256"; do set -- $cfg; ASM_B200_LANES=$1 ASM_B200_CHUNK_MB=$2 python tools/quick_bench.py 2048 32 1 2; done
echo "== 4096 unpadded, 128 MB per sample"
for cfg in "3 216" "1 128" "2 256"; do set -- $cfg; ASM_B200_LANES=$1 ASM_B200_CHUNK_MB=$2 python tools/quick_bench.py 4096 16 0 2; done
} > gpurun_out/r2ac_sweep.log 2>&1

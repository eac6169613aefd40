#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
for v in v0 nd v0 nd; do echo "== $v (nd: no L2 discards)"; ASM_B200_LIB=$D/libasm_b200_$v.so python tools/quick_bench.py 1024 512 0 10; ASM_B200_LIB=$D/libasm_b200_$v.so python tools/quick_bench.py 512 1024 1 10; done
echo "== nd with 3 lanes x 216 MB"; ASM_B200_LIB=$D/libasm_b200_nd.so ASM_B200_LANES=3 ASM_B200_CHUNK_MB=216 python tools/quick_bench.py 1024 512 0 10
} > gpurun_out/r2ak.log 2>&1
ASM_B200_LIB=$D/libasm_b200_nd.so ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/r2ak_dram_nd.csv python tools/prof_case.py 1024 108 0 1 > /dev/null 2>&1

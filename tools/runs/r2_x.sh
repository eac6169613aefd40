#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
export ASM_B200_LIB=$D/libasm_b200_tune.so
ncu --set full --clock-control none --import-source on -k regex:k_cluster256 -c 1 -o gpurun_out/r2x_c256 -f python tools/prof_case.py 256 592 0 1 1 > gpurun_out/r2x_ncu.log 2>&1

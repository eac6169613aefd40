#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
{
echo "== product lib (6-warp, cross-proxy fence before the next request)"; python tools/diag_energy.py 1024 48 25 | cut -c1-160
echo "== 2048"; python tools/diag_energy.py 2048 12 8 | cut -c1-160
echo "== 512 pad"; python tools/diag_energy.py 512 96 8 | cut -c1-160
echo "== perf"; python tools/quick_bench.py 1024 512; python tools/quick_bench.py 2048 128
} > gpurun_out/r2g_diag3.log 2>&1

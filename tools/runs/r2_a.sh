#!/bin/bash
# round-2 first GPU batch: baseline tests + knob sweep with the round-1 build
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/r2a_tests.log
{
echo "== default"; python tools/quick_bench.py 1024 512
echo "== pass times default"; python tools/pass_times.py 1024 108
for l in 2 3 4; do for c in 1 2; do
  echo "== BULK=0 LANES=$l CTAS_PER_SM=$c"; ASM_B200_BULK=0 ASM_B200_LANES=$l ASM_B200_CTAS_PER_SM=$c python tools/quick_bench.py 1024 512
done; done
for l in 3 4; do
  echo "== BULK=3 LANES=$l CTAS_PER_SM=1"; ASM_B200_LANES=$l ASM_B200_CTAS_PER_SM=1 python tools/quick_bench.py 1024 512
done
for mb in 96 144 288; do
  echo "== CHUNK_MB=$mb"; ASM_B200_CHUNK_MB=$mb python tools/quick_bench.py 1024 512
done
echo "== other shapes"
python tools/quick_bench.py 256 4096
python tools/quick_bench.py 256 4096 1
python tools/quick_bench.py 512 1024
python tools/quick_bench.py 512 1024 1
python tools/quick_bench.py 1024 128 1
python tools/quick_bench.py 2048 128
python tools/quick_bench.py 2048 32 1
python tools/quick_bench.py 128 8192 1
python tools/pass_times.py 256 1024
python tools/pass_times.py 1024 32 1
python tools/pass_times.py 2048 32
} > gpurun_out/r2a_sweep.log 2>&1

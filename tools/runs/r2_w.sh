#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
export ASM_B200_LIB=$D/libasm_b200_tune.so
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_paths.py -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2w_tests.log
{
for c in 1 0; do
  echo "== CLUSTER256=$c"
  ASM_B200_CLUSTER256=$c python bench.py --config c2 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['roofline']['frac'], d['parity'], d['gpu_launches'])"
  ASM_B200_CLUSTER256=$c python tools/quick_bench.py 256 4096 0 10
  ASM_B200_CLUSTER256=$c python tools/quick_bench.py 128 8192 1 10
done
} > gpurun_out/r2w_c2.log 2>&1

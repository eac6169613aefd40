#!/bin/bash
mkdir -p gpurun_out
D=$PWD/style_transfer_based_holographic_imaging_b200
export ASM_B200_LIB=$D/libasm_b200_tune.so
{
echo "== c2 (256^2 forward only, B = 4096): lanes x MB in flight"
for l in 2 3 4 6; do for mb in 24 48 72 96 144; do
  export ASM_B200_LANES=$l ASM_B200_CHUNK_MB=$mb; python bench.py --config c2 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline --no-e2e --no-parity | python -c "import json,sys,os; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(os.environ['ASM_B200_LANES'], os.environ['ASM_B200_CHUNK_MB'], round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'])"
done; done
unset ASM_B200_LANES ASM_B200_CHUNK_MB
echo "== 128^2 padded (MNIST demo shape), B = 8192, fwd+adj"
for cfg in "3 24" "3 72"; do set -- $cfg; ASM_B200_LANES=$1 ASM_B200_CHUNK_MB=$2 python tools/quick_bench.py 128 8192 1 10; done
} > gpurun_out/r2ag_c2.log 2>&1

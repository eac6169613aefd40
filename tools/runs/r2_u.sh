#!/bin/bash
# 8-GPU batch: strong scaling of c3 (global batch 512 split over the GPUs) at N = 8 and 4, and the DDP training step on 8 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong --no-cpu --no-gpu-baseline > gpurun_out/r2u_bench_c3_strong_n8.json 2> gpurun_out/r2u.err
$TR --nproc-per-node 4 --master-port 29622 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong --no-cpu --no-gpu-baseline --no-e2e > gpurun_out/r2u_bench_c3_strong_n4.json 2>> gpurun_out/r2u.err
$TR --nproc-per-node 8 --master-port 29623 examples/train_step.py --steps 10 --out gpurun_out/r2u_train_n8.json > gpurun_out/r2u_train_n8.log 2>&1
nvidia-smi topo -m > gpurun_out/r2u_topo.log 2>&1

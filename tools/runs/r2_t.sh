#!/bin/bash
# 2-GPU batch: unwrap tests on one GPU, then strong / weak scaling lines and the DDP training step on 2 GPUs
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_unwrap.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2t_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-cpu --no-gpu-baseline > gpurun_out/r2t_bench_c3_strong_n2.json 2> gpurun_out/r2t_n2.err
$TR --nproc-per-node 2 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/r2t_bench_c3_weak_n2.json 2>> gpurun_out/r2t_n2.err


$TR --nproc-per-node 2 --master-port 29613 examples/train_step.py --steps 10 --out gpurun_out/r2t_train_n2.json > gpurun_out/r2t_train_n2.log 2>&1

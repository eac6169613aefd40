#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python bench.py --bw-test > gpurun_out/r2al_bw_n1.json 2> gpurun_out/r2al.err
$TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --bw-test > gpurun_out/r2al_bw_n8.json 2>> gpurun_out/r2al.err

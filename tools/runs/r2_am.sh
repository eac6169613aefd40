#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2 --master-port 29641 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2am_n2.json 2> gpurun_out/r2am.err

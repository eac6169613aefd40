#!/bin/bash
# round-2 re-entry batch: full GPU test suite, then the measurement batch (bench lines of every config, launch list, ncu captures, DRAM traffic)
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/r2h_tests.log
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2h_gpu.log
bash tools/runs/r2_profiles.sh

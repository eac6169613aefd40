#!/bin/bash
mkdir -p gpurun_out
for c in c4pad c3pad; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done

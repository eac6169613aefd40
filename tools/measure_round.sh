set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -c 600 gpurun_out/bench_r1c.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1c.json 2>/dev/null; tail -c 300 gpurun_out/bench_ref_r1c.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bench_r1c.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/dram_r1c.csv python tools/prof_case.py 1024 108 0 1 > gpurun_out/dram_r1c.log 2>&1
ASM_B200_FLOW=1 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/dram_flow_r1c.csv python tools/prof_case.py 1024 108 0 1 > gpurun_out/dram_flow_r1c.log 2>&1
ASM_B200_FLOW=1 ASM_B200_RING=6 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/dram_flow6_r1c.csv python tools/prof_case.py 1024 108 0 1 > gpurun_out/dram_flow6_r1c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k32 -c 7 -o gpurun_out/prof_r1c -f python tools/prof_case.py 1024 27 0 1 > gpurun_out/ncu_r1c.log 2>&1
du -sh gpurun_out

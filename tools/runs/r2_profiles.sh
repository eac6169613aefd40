#!/bin/bash
# round-2 measurement batch: bench lines for every config, launch list, ncu --set full of the default kernels, DRAM traffic
mkdir -p gpurun_out
for c in c3 c2 c3pad c4 c4pad; do
  python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done
python bench.py --config c3 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_reference.json 2>> gpurun_out/r2_bench_c3.err
python bench.py --config c3 --impl reference-cuda --steps 3 --warmup 1 > gpurun_out/r2_bench_c3_reference_cuda.json 2>> gpurun_out/r2_bench_c3.err
# launch list of the default bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --batch 64 --no-e2e --no-cpu --no-gpu-baseline --no-parity > gpurun_out/r2_ncu_launches.log 2>&1
# full captures of the default kernels: FFT 1024 (k32t), FFT 256 (generic 16-point kernels), FFT 2048 (k64)
EXTRA=lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_write_bytes.sum,smsp__inst_executed_pipe_lsu.sum,l1tex__m_xbar2l1tex_read_bytes.sum
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k32t -c 6 -o gpurun_out/r2_k32t -f python tools/prof_case.py 1024 27 0 1 > gpurun_out/r2_ncu_k32t.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:"k_rows_fwd|k_cols|k_rows_inv" -c 3 -o gpurun_out/r2_n8 -f python tools/prof_case.py 256 1184 0 1 1 > gpurun_out/r2_ncu_n8.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k64 -c 4 -o gpurun_out/r2_k64 -f python tools/prof_case.py 2048 6 0 1 1 > gpurun_out/r2_ncu_k64.log 2>&1
# DRAM traffic of whole calls (no cache control, application replay)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/r2_dram_c3.csv python tools/prof_case.py 1024 108 0 1 > gpurun_out/r2_dram_c3.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/r2_dram_c2.csv python tools/prof_case.py 256 2368 0 1 1 > gpurun_out/r2_dram_c2.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --replay-mode application --csv --log-file gpurun_out/r2_dram_c4.csv python tools/prof_case.py 2048 32 0 1 > gpurun_out/r2_dram_c4.log 2>&1
# gpurun copies at most 64 MiB back: keep the FFT-1024 report, summarise the other two on the box
python tools/ncu_summary.py gpurun_out/r2_n8.ncu-rep gpurun_out/r2_ncu_fft256_raw.md > /dev/null 2>&1 && rm -f gpurun_out/r2_n8.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_k64.ncu-rep gpurun_out/r2_ncu_k64_raw.md > /dev/null 2>&1 && rm -f gpurun_out/r2_k64.ncu-rep
du -sh gpurun_out

#!/bin/bash
mkdir -p gpurun_out
EXTRA=lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_write_bytes.sum,smsp__inst_executed_pipe_lsu.sum,l1tex__m_xbar2l1tex_read_bytes.sum
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:"k_rows_fwd|k_cols|k_rows_inv" -c 3 -o gpurun_out/r2_n8 -f python tools/prof_case.py 256 1184 0 1 1 > gpurun_out/r2_ncu_n8.log 2>&1
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k64 -c 4 -o gpurun_out/r2_k64 -f python tools/prof_case.py 2048 6 0 1 1 > gpurun_out/r2_ncu_k64.log 2>&1

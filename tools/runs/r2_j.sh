#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2j_tests.log
EXTRA=lts__t_bytes.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_l1tex2xbar_write_bytes.sum,smsp__inst_executed_pipe_lsu.sum,l1tex__m_xbar2l1tex_read_bytes.sum
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:k32t -c 6 -o gpurun_out/r2j_k32t -f python tools/prof_case.py 1024 27 0 1 > gpurun_out/r2j_ncu.log 2>&1

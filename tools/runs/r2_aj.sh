#!/bin/bash
mkdir -p gpurun_out
ncu --metrics lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors.sum,gpu__time_duration.sum --cache-control none -k regex:vectorized -c 3 --csv --log-file gpurun_out/copy_probe.csv python tools/copy_probe.py > /dev/null 2>&1

#!/bin/bash
# ncu capture of the standalone NTT kernels (config 2 microbenchmark), 1 GPU, under gpurun
TAG=${1:-ntt}
CMD="python scripts/quick_bench.py 1024"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_ntt' -s 4 -c 4 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log

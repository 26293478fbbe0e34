#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md applied to bench.py (run under gpurun, 1 GPU).
# usage: scripts/gpu_profile.sh <tag>     -> gpurun_out/<tag>_{launches.csv,prof.ncu-rep,...}
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --batch 1184 --no-cpu-baseline --no-e2e --no-configs"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_ext_conv|k_ext_ntt|k_tensor_intt|k_tensor_floor|k_floor_sk|k_digit_ntt|k_ks_intt|k_ks_finish|k_ks_tail' -s 15 -c 5 -o $OUT/${TAG}_prof -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
tail -3 $OUT/${TAG}_ncu_full.log

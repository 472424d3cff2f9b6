#!/usr/bin/env bash
# Round-1 profile refresh: launch lists + full captures (each ncu run follows a plain run of the same command that exited 0).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 1024 --no-cpu-baseline"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_prove.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"4, dr::CommitBody" -s 2 -c 2 -o gpurun_out/prof_commit3 $CMD > gpurun_out/ncu_full.log 2>&1
CMD2="python tools/msm_one.py 20"
$CMD2 > gpurun_out/plain_msm.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_msm20.csv $CMD2 > gpurun_out/ncu_list_msm.log 2>&1
$CMD2 > gpurun_out/plain_msm2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"MsmUnitSumBody" -s 1 -c 1 -o gpurun_out/prof_msm_unit $CMD2 > gpurun_out/ncu_full_msm.log 2>&1
tail -n 2 gpurun_out/plain_msm.log gpurun_out/ncu_full.log gpurun_out/ncu_full_msm.log

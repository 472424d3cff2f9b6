#!/usr/bin/env bash
# Round-2 first GPU visit: tests, bench lines (strong 4096, the 512-proof shard of an 8-GPU run), launch list, ncu captures of the
# kernels the round-1 verdict found without evidence.  Each ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gputests.log 2>&1; echo "gputests rc=$?"
tail -n 3 gpurun_out/r2a_gputests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
python bench.py --steps 8 --warmup 3 --total 512 --saturated-batch 0 --no-cpu-baseline > gpurun_out/r2a_bench_n1_total512.json 2> gpurun_out/r2a_bench_512.err; echo "bench512 rc=$?"
python bench.py --steps 5 --warmup 3 --total 1024 --saturated-batch 0 --no-cpu-baseline > gpurun_out/r2a_bench_n1_total1024.json 2>> gpurun_out/r2a_bench_512.err
python tools/gpu_probe_r2.py 1024 20000 10 > gpurun_out/r2a_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2a_probe.log
CMD="python tools/gpu_probe_r2.py 512 4096 10"
$CMD > gpurun_out/r2a_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_probe512.csv $CMD > gpurun_out/r2a_ncu_list.log 2>&1
$CMD > gpurun_out/r2a_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"PedersenProveBody|WitnessBody|WitnessLdeBody|ConstraintBody|QuotientInttBody|WitnessInttBody|OpenQuotientsBody|EvalBody" -c 8 -o gpurun_out/r2a_prof_prove_kernels $CMD > gpurun_out/r2a_ncu_full.log 2>&1
tail -n 2 gpurun_out/r2a_ncu_full.log
$CMD > gpurun_out/r2a_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"IetfVerifyBody|PedersenVerifyBody|TeDecodeManyBody" -s 6 -c 5 -o gpurun_out/r2a_prof_verify_kernels $CMD > gpurun_out/r2a_ncu_full2.log 2>&1
tail -n 2 gpurun_out/r2a_ncu_full2.log
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv

#!/usr/bin/env bash
# Final single-GPU evidence of round 2: tests, the bench line, the launch list of the same command, one full capture of the dominant kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2z
python -m pytest tests -m gpu -q --tb=short > ${O}_gputests.log 2>&1; echo "gputests rc=$?"; tail -n 2 ${O}_gputests.log
python bench.py --steps 10 --warmup 4 > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('${O}_bench_n1.json')); print({k: d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['phase_ms_per_step'], d.get('saturated'), d.get('cpu_baseline'), d['config']['parity'])"
python bench.py --impl reference --steps 4 --warmup 1 > ${O}_bench_reference_arm.json 2> ${O}_bench_reference_arm.err; echo "ref rc=$?"; cut -c1-300 ${O}_bench_reference_arm.json
CMD="python bench.py --steps 1 --warmup 1 --saturated-batch 0 --no-cpu-baseline"
$CMD > ${O}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file ${O}_launches_bench4096.csv $CMD > ${O}_ncu_list.log 2>&1
python tools/summarize_launches.py ${O}_launches_bench4096.csv TableBuildBody G1NttStageBody G1PrefixSumBody G1ScaleBody > ${O}_launch_share_bench4096.txt 2>&1; head -n 12 ${O}_launch_share_bench4096.txt
CMD3="python bench.py --steps 1 --warmup 1 --total 1024 --saturated-batch 0 --no-cpu-baseline"
$CMD3 > ${O}_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"CommitBodyT" -s 4 -c 1 -o ${O}_prof_commit $CMD3 > ${O}_ncu_full3.log 2>&1
python tools/ncu_extract.py ${O}_prof_commit.ncu-rep ${O}_prof_commit.csv > ${O}_prof_commit.txt 2>&1; cat ${O}_prof_commit.txt; rm -f ${O}_prof_commit.ncu-rep
python tools/gpu_microbench_r2.py > ${O}_microbench.json 2>&1
du -sh gpurun_out

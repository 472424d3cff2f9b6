#!/usr/bin/env bash
# Round-2 GPU visit: tests, bench lines, probe, launch list, ncu captures.  Usage: bash tools/gpu_r2_run.sh <tag> [steps...]
# steps: tests bench bench512 probe micro commit_ab list ncu_prove ncu_verify ncu_commit.  ncu reports are reduced to CSV on the box (tools/ncu_extract.py)
# because gpurun only brings back 64 MiB; each ncu run follows a plain run of the same command that exited 0.
set -u
TAG=$1; shift
STEPS=" $* "
mkdir -p gpurun_out
O=gpurun_out/${TAG}
python -c "import ctypes, __graft_entry__ as g; lib = ctypes.CDLL(\"dot_ring_b200/libdotring_b200.so\"); m = [s for s in g.declared_symbols() if not hasattr(lib, s)]; assert not m, m" || exit 1
has() { [[ "$STEPS" == *" $1 "* ]]; }
if has tests; then
  python -m pytest tests -m gpu -q --tb=short > ${O}_gputests.log 2>&1; echo "gputests rc=$?"
  tail -n 40 ${O}_gputests.log
fi
if has bench; then
  python bench.py --steps 5 --warmup 3 > ${O}_bench_n1.json 2> ${O}_bench_n1.err; echo "bench rc=$?"; tail -c 600 ${O}_bench_n1.err
  python -c "import json,sys; d=json.load(open('${O}_bench_n1.json')); print({k: d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['frac'], d['phase_ms_per_step'], d.get('saturated'))"
fi
if has bench512; then
  for T in 512 1024 2048; do
    python bench.py --steps 6 --warmup 3 --total $T --saturated-batch 0 --no-cpu-baseline > ${O}_bench_n1_total$T.json 2>> ${O}_bench_small.err; echo "bench$T rc=$?"
    python -c "import json,sys; d=json.load(open('${O}_bench_n1_total$T.json')); print($T, {k: d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['phase_ms_per_step'])"
  done
fi
if has probe; then
  python tools/gpu_probe_r2.py 1024 ${PROBE_ITEMS:-100000} 10 > ${O}_probe.log 2>&1; echo "probe rc=$?"; cat ${O}_probe.log
fi
if has timeline; then
  python tools/gpu_probe_timeline.py 10 > ${O}_timeline.log 2>&1; echo "timeline rc=$?"; grep -v "^\[dot_ring_b200\] prove: free" ${O}_timeline.log | tail -n 80
fi
if has micro; then
  python tools/gpu_microbench_r2.py > ${O}_microbench.json 2>&1; cat ${O}_microbench.json
fi
if has commit_ab; then
  python tools/gpu_commit_ab.py default build/var/lib_m3.so build/var/lib_t96m5.so build/var/lib_t64m8.so > ${O}_commit_ab.log 2>&1; cat ${O}_commit_ab.log
fi
CMD="python tools/gpu_probe_r2.py 512 4096 10"
if has list; then
  $CMD > ${O}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file ${O}_launches_probe512.csv $CMD > ${O}_ncu_list.log 2>&1
fi
extract() { python tools/ncu_extract.py $1.ncu-rep $1.csv > $1.txt 2>&1; ls -la $1.ncu-rep; rm -f $1.ncu-rep; }
if has ncu_prove; then
  $CMD > ${O}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --kernel-name-base demangled -k regex:"PedersenStartBody|PedersenFinishBody|WitnessBody|WitnessLdeBody|ConstraintBody|QuotientInttBody|WitnessInttBody|OpenQuotientsBody|EvalBody" -c 9 -o ${O}_prof_prove_kernels $CMD > ${O}_ncu_full.log 2>&1
  extract ${O}_prof_prove_kernels; cat ${O}_prof_prove_kernels.txt
fi
if has ncu_verify; then
  $CMD > ${O}_plain3.log 2>&1 &&
  ncu --set full --clock-control none --kernel-name-base demangled -k regex:"IetfVerifyBody|PedersenVerifyBody|TeDecodeManyBody" -s 6 -c 5 -o ${O}_prof_verify_kernels $CMD > ${O}_ncu_full2.log 2>&1
  extract ${O}_prof_verify_kernels; cat ${O}_prof_verify_kernels.txt
fi
if has ncu_start; then
  $CMD > ${O}_plain5.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"PedersenStartBody" -s 1 -c 1 -o ${O}_prof_start $CMD > ${O}_ncu_full5.log 2>&1
  python tools/ncu_extract.py ${O}_prof_start.ncu-rep ${O}_prof_start.csv > ${O}_prof_start.txt 2>&1; cat ${O}_prof_start.txt; ls -la ${O}_prof_start.ncu-rep
fi
if has ncu_commit; then
  CMD3="python bench.py --steps 1 --warmup 1 --total 1024 --saturated-batch 0 --no-cpu-baseline"
  $CMD3 > ${O}_plain4.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"CommitBodyT" -s 4 -c 1 -o ${O}_prof_commit $CMD3 > ${O}_ncu_full3.log 2>&1
  python tools/ncu_extract.py ${O}_prof_commit.ncu-rep ${O}_prof_commit.csv > ${O}_prof_commit.txt 2>&1; cat ${O}_prof_commit.txt; ls -la ${O}_prof_commit.ncu-rep
fi
du -sh gpurun_out

#!/usr/bin/env bash
# A/B: same sources, different build options (variants/*.so are built locally, not committed)
set -u
mkdir -p gpurun_out
for v in ${VARIANTS:-default}; do
  if [ $v = default ]; then unset DR_LIB; else export DR_LIB=$PWD/variants/lib_$v.so; fi
  for rep in 1 2; do
  PROBE_OUT=gpurun_out/ab_probe_$v.json python tools/gpu_probe.py > gpurun_out/ab_probe_$v.log 2>&1
  echo "== $v"; grep -E "g1_madd|'batch': 4096|'n': 6145, 'batch': 1024" gpurun_out/ab_probe_$v.log
  done
  if [ "${WITH_PROVE:-0}" = 1 ]; then
  PROVE_NS=4096,4096 PROBE_OUT=gpurun_out/ab_prove_$v.json python tools/gpu_probe_prove.py > gpurun_out/ab_prove_$v.log 2>&1
  tail -n 2 gpurun_out/ab_prove_$v.log
  fi
done

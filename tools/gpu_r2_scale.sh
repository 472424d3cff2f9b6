#!/usr/bin/env bash
# Strong scaling of BASELINE configs[1] (4096 proofs per step) on one 8-GPU box: torchrun ranks (the driver's launch) at N = 8, 4, 2 and the
# in-library dispatcher (one process, EnginePool) at N = 8.  Usage: bash tools/gpu_r2_scale.sh <tag>
set -u
TAG=$1
mkdir -p gpurun_out
run() { # N
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 6 --warmup 3 \
    > gpurun_out/${TAG}_bench_n$1.json 2> gpurun_out/${TAG}_bench_n$1.err
  echo "N=$1 rc=$?"; python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n$1.json')); print({k: d[k] for k in ('n_gpus','value','ms_per_step','scaling')}, d['e2e']['value'], d.get('saturated'))"
}
nvidia-smi --query-gpu=index,name,memory.used --format=csv | head -12
run 8
python bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/${TAG}_bench_n8_pool.json 2> gpurun_out/${TAG}_bench_n8_pool.err; echo "pool rc=$?"
python -c "import json; d=json.load(open('gpurun_out/${TAG}_bench_n8_pool.json')); print('pool', {k: d[k] for k in ('n_gpus','value','ms_per_step','scaling')}, d['e2e']['value'], d['config']['launch'], d.get('saturated'))"
run 4
run 2
tail -n 3 gpurun_out/${TAG}_bench_n8.err gpurun_out/${TAG}_bench_n8_pool.err

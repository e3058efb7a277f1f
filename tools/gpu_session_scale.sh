#!/bin/bash
# One bench.py run at N ranks with the default transport (what the driver's scaling run does).  usage: gpu_session_scale.sh <tag> <N> [steps]
tag=${1:-s}; N=${2:-8}; steps=${3:-5}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu_$tag.txt 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps $steps --warmup 3 > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err
echo "bench N=$N rc=$?"
tail -c 1500 gpurun_out/bench_${tag}_n$N.json
tail -5 gpurun_out/bench_${tag}_n$N.err
true

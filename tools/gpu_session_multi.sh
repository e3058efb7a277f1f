#!/bin/bash
# Multi-GPU session (gpurun --gpus N): two-rank hardware parity test + bench at N ranks with both transports.
tag=${1:-m}; N=${2:-2}; steps=${3:-6}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu_$tag.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout 500 > gpurun_out/pytest_multi_$tag.log 2>&1; echo "pytest multi rc=$?"
timeout 300 python -m pytest tests/test_gpu_model.py -m gpu -q -k zero_shot > gpurun_out/pytest_zs_$tag.log 2>&1; echo "zero-shot rc=$?"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps $steps --warmup 3 > gpurun_out/bench_${tag}_$name.json 2> gpurun_out/bench_${tag}_$name.err
  echo "bench $name rc=$?"
}
run peer VL_DUMMY=1
run nccl VL_COMM=nccl
tail -4 gpurun_out/pytest_multi_$tag.log
tail -3 gpurun_out/bench_${tag}_peer.err
true

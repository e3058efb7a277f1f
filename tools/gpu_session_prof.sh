#!/bin/bash
# Profiling session: ncu --set full (with source) of the kernels VERDICT names + the launch list of one bench step.
tag=${1:-p}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:gemm2 -s 3 -c 1 -f -o gpurun_out/${tag}_gelu python tools/ncu_one.py gelu > gpurun_out/${tag}_gelu.log 2>&1; echo "gelu rc=$?"
timeout 300 $NCU -k regex:gemm2 -s 3 -c 1 -f -o gpurun_out/${tag}_gelu_bwd python tools/ncu_one.py gelu_bwd > gpurun_out/${tag}_gelu_bwd.log 2>&1; echo "gelu_bwd rc=$?"
timeout 300 $NCU -k regex:attn_bwd3 -s 2 -c 1 -f -o gpurun_out/${tag}_attn_bwd python tools/ncu_one.py attn_bwd > gpurun_out/${tag}_attn_bwd.log 2>&1; echo "attn_bwd rc=$?"
timeout 300 $NCU -k regex:attn_fwd2 -s 2 -c 1 -f -o gpurun_out/${tag}_attn_fwd python tools/ncu_one.py attn_fwd > gpurun_out/${tag}_attn_fwd.log 2>&1; echo "attn_fwd rc=$?"
timeout 300 $NCU -k regex:layernorm_bwd -s 2 -c 1 -f -o gpurun_out/${tag}_ln_bwd python tools/run_ln_once.py > gpurun_out/${tag}_ln_bwd.log 2>&1; echo "ln bwd rc=$?"
timeout 300 $NCU -k regex:layernorm_fwd -s 2 -c 1 -f -o gpurun_out/${tag}_ln_fwd python tools/run_ln_once.py > gpurun_out/${tag}_ln_fwd.log 2>&1; echo "ln fwd rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 900 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-torch-baseline > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/${tag}_* | head -20
true

#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/probe_gelu_gemm.py > gpurun_out/probe_gelu.log 2>&1; echo "probe rc=$?"; cat gpurun_out/probe_gelu.log | grep gelu-gemm
VL_LIB_PATH= timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm or geglu" --timeout 200 > gpurun_out/pytest_gelu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gelu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "layernorm or row_kernels" --timeout 120 > gpurun_out/pytest_ln.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ln.log
timeout 120 python tools/probe_rows.py > gpurun_out/probe_rows.log 2>&1; echo "probe rc=$?"; grep -i "layernorm\|ln" gpurun_out/probe_rows.log | head -12

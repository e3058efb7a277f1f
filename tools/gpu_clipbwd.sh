#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_contrastive_backward or contrastive_epilogues" --timeout 120 > gpurun_out/pytest_clipbwd.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/pytest_clipbwd.log

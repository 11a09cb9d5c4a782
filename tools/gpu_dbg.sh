#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -s 2>&1 | grep -v "^$" | grep -i "shape\|differ\|row\|passed\|failed" | head -80 > gpurun_out/dbg_gemm.log
timeout 600 python -m pytest tests/test_gpu_nets.py -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed" | tail -12 > gpurun_out/dbg_nets.log
cat gpurun_out/dbg_gemm.log gpurun_out/dbg_nets.log

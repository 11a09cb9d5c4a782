#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed\|Error" | tail -12 | tee gpurun_out/dbg_pipe.log
timeout 400 /usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 8 -f -o gpurun_out/final_gemm python bench.py --steps 2 --warmup 1 --serial --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/final_gemm.ncu-rep

#!/bin/bash
# round 2, call A: the new B=256 / config #1 parity tests, the whole GPU suite, a launch list and one --set full capture
# of every kernel of a serial step (raw counters for profiles/)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_batch256.py -q -m gpu -s -x 2>&1 | tail -25 > $out/r2a_b256.log; tail -8 $out/r2a_b256.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/r2a_tests.log; tail -3 $out/r2a_tests.log
NCU=/usr/local/cuda/bin/ncu
B="python bench.py --steps 1 --warmup 1 --serial --no-cpu-baseline"
timeout 600 $NCU --set full --clock-control none -s 52 -c 60 -f -o $out/r2a_step_full $B > $out/r2a_ncu_full.log 2>&1
ls -la $out/r2a_* | awk '{print $5, $9}'

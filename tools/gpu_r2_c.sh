#!/bin/bash
# round 2, call C: post-path / stream / pipeline tests again + variant B (WeSpeaker ResNet34) bring-up
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_zz_wespeaker.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -60 > $out/r2c_wespeaker.log; tail -45 $out/r2c_wespeaker.log
timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_stream.py tests/test_gpu_pipeline.py tests/test_gpu_nets.py -q -m gpu 2>&1 | tail -30 > $out/r2c_tests.log; tail -15 $out/r2c_tests.log

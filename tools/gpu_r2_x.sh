#!/bin/bash
# round 2, call X: which side moves?  md5 of the CPU oracle output and of the CUDA output in a series of processes after the suite
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -2
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  timeout 300 python tests/manual/after_suite_probe.py "p$i" 2>&1 | grep -E "^p$i" | cut -c1-520
done | tee $out/r2x_probe.log
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA node\(s\)" ; env | grep -i -E "dnnl|mkl|omp|onednn" | head

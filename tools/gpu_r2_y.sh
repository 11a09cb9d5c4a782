#!/bin/bash
# round 2, call Y (CPU work on the GPU box): reproducibility of the torch CPU float32 oracle across processes, with and without oneDNN
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
for i in $(seq 1 18); do timeout 120 python tests/manual/oracle_repro_probe.py "default-$i" 2>&1 | tail -1; done | tee $out/r2y_default.log | awk '{print $5}' | sort | uniq -c
for i in $(seq 1 18); do ORACLE_MKLDNN=0 timeout 120 python tests/manual/oracle_repro_probe.py "nomkldnn-$i" 2>&1 | tail -1; done | tee $out/r2y_nomkldnn.log | awk '{print $5}' | sort | uniq -c
for i in $(seq 1 12); do ORACLE_THREADS=16 timeout 120 python tests/manual/oracle_repro_probe.py "t16-$i" 2>&1 | tail -1; done | tee $out/r2y_t16.log | awk '{print $5}' | sort | uniq -c
grep -h -v "c5cbeb84 c5cbeb84 c5cbeb84" $out/r2y_default.log $out/r2y_t16.log | head -8
head -2 $out/r2y_nomkldnn.log

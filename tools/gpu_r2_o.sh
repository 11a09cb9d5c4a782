#!/bin/bash
# round 2, call O: which bench leg makes the parity check of the xvector configuration fail (seen once in call N)?
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
export DG_BENCH_PARITY_DETAIL=1
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 6 --warmup 3 --no-cpu-baseline $FLAGS > $out/r2o_$name.json 2> $out/r2o_$name.err; echo "== $name rc=$?"; grep -E "parity detail|parity_failed|fails its parity" $out/r2o_$name.err | head -5; }
FLAGS="--no-stream-leg --no-pipeline-call" run none X=1
FLAGS="--no-pipeline-call" run stream X=1
FLAGS="--no-stream-leg" run call X=1
FLAGS="" run both X=1
FLAGS="" run both_nopool3 DG_NO_POOL3_FUSE=1
FLAGS="" run both_rows16 DG_LSTM_ROWS=16
FLAGS="" run both_again X=1
DG_CLUSTER_TIMING=1 timeout 120 python tests/manual/cluster_timing.py 2>&1 | tail -6

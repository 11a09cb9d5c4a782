#!/bin/bash
# round 2, call Q: the parity check of bench.py --steps 20 failed twice (calls N, P) and never at --steps 6 (call O): bisect
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
export DG_BENCH_PARITY_DETAIL=1
run() { name=$1; shift; timeout 300 env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline $FLAGS > $out/r2q_$name.json 2> $out/r2q_$name.err; echo "== $name rc=$?"; grep -E "parity detail|parity_failed" $out/r2q_$name.err | cut -c1-330 | head -5; }
FLAGS="" run default1 X=1
FLAGS="" run default2 X=1
FLAGS="" run nopool3 DG_NO_POOL3_FUSE=1
FLAGS="" run rows16 DG_LSTM_ROWS=16
FLAGS="--no-pipeline-call" run nocall X=1
FLAGS="--no-stream-leg" run nostream X=1
FLAGS="--no-stream-leg --no-pipeline-call" run neither X=1
FLAGS="" run nodedup DG_CALL_NO_DEDUP=1
FLAGS="" run nowscale DG_NO_WSCALE=1

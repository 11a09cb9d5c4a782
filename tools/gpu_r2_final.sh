#!/bin/bash
# round 2, last call: lstm timing, then the driver's sequence on HEAD (suite -> smoke -> bench with default arguments -> reference arm)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
DG_LSTM_ROWS=16 DG_LSTM_CELLS=4 timeout 300 python tests/manual/lstm_timing.py 256 2>&1 | grep -E "lstm_rec timing|seg max" | tail -2 | tee $out/r2final_lstm_timing.log
bash tools/gpu_r2_v.sh

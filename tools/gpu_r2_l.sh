#!/bin/bash
# round 2, call L: LSTM recurrence with three completion points; 16 vs 8 rows per CTA; phase timing; __call__ phases
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
for shape in "16 4" "16 8" "8 2" "8 4"; do
  set -- $shape
  DG_LSTM_ROWS=$1 DG_LSTM_CELLS=$2 timeout 300 python tests/manual/lstm_timing.py 256 2>&1 | grep -E "lstm_rec timing|seg max|Error|error" | head -4
done > $out/r2l_lstm_timing.log 2>&1
cat $out/r2l_lstm_timing.log
DG_LSTM_ROWS=8 timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_shapes.py tests/test_gpu_batch256.py -q -m gpu -x 2>&1 | tail -5 > $out/r2l_tests_rows8.log; tail -3 $out/r2l_tests_rows8.log
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_shapes.py tests/test_gpu_batch256.py -q -m gpu -x 2>&1 | tail -5 > $out/r2l_tests_rows16.log; tail -3 $out/r2l_tests_rows16.log
DG_CALL_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg > $out/r2l_bench_rows16.json 2> $out/r2l_bench_rows16.err; grep dg_pipeline_call_host $out/r2l_bench_rows16.err | tail -2
DG_CALL_TIMING=1 DG_LSTM_ROWS=8 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg > $out/r2l_bench_rows8.json 2> $out/r2l_bench_rows8.err; grep dg_pipeline_call_host $out/r2l_bench_rows8.err | tail -2
DG_LSTM_ROWS=8 DG_LSTM_CELLS=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2l_bench_rows8c4.json 2> $out/r2l_bench_rows8c4.err
python - <<'PY'
import json
for f in ("r2l_bench_rows16.json", "r2l_bench_rows8.json", "r2l_bench_rows8c4.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "call", d.get("e2e_pipeline_call"), "parity", d.get("parity"))
        print(" rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

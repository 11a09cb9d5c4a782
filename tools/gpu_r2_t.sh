#!/bin/bash
# round 2, call T: gate pre-activations fetched one step ahead in lstm_rec; clocks sampled from the warm-up on
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
for shape in "16 4" "8 2"; do
  set -- $shape
  DG_LSTM_ROWS=$1 DG_LSTM_CELLS=$2 timeout 300 python tests/manual/lstm_timing.py 256 2>&1 | grep -E "lstm_rec timing|seg max|Error|error" | tail -2
done > $out/r2t_lstm_timing.log 2>&1
cat $out/r2t_lstm_timing.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $out/r2t_tests.log; tail -3 $out/r2t_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2t_bench.json 2> $out/r2t_bench.err; echo "bench rc=$?"; grep -E "parity_failed" $out/r2t_bench.err | cut -c1-600
DG_NO_TMA_STORE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2t_bench_notmastore.json 2> $out/r2t_bench_notmastore.err
python - <<'PY'
import json
for f in ["r2t_bench", "r2t_bench_notmastore"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("e2e_pipeline_call") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), "parity", d.get("parity"), d.get("clocks"))
        print("  rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

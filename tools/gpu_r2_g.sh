#!/bin/bash
# round 2, call G: TC_POOL v2 fixed; LSTM 16 cell warps vs 8; serial (kernel-alone) timings
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/r2g_tests.log; tail -8 $out/r2g_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2g_bench.json 2> $out/r2g_bench.err; tail -3 $out/r2g_bench.err
DG_LSTM_CELLS=8 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2g_bench_cells8.json 2> $out/r2g_bench_cells8.err
DG_NO_POOL_FUSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2g_bench_nofuse.json 2> $out/r2g_bench_nofuse.err
timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2g_bench_serial.json 2> $out/r2g_bench_serial.err
DG_NO_POOL_FUSE=1 DG_LSTM_CELLS=8 timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2g_bench_serial_old.json 2> $out/r2g_bench_serial_old.err
python - <<'PY'
import json
for f in ("r2g_bench.json", "r2g_bench_cells8.json", "r2g_bench_nofuse.json", "r2g_bench_serial.json", "r2g_bench_serial_old.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", (d.get("e2e_pipeline_call") or {}).get("ms_per_call"), "parity", d.get("parity"))
        print(d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

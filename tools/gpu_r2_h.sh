#!/bin/bash
# round 2, call H: stream-major pool passes, unrolled TC_POOL accumulation
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests/test_gpu_nets.py tests/test_gpu_pipeline.py tests/test_gpu_shapes.py tests/test_gpu_batch256.py tests/test_zz_stream_form.py tests/test_gpu_stream.py -q -m gpu 2>&1 | tail -15 > $out/r2h_tests.log; tail -8 $out/r2h_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2h_bench.json 2> $out/r2h_bench.err; tail -3 $out/r2h_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2h_bench_serial.json 2> $out/r2h_bench_serial.err
python - <<'PY'
import json
for f in ("r2h_bench.json", "r2h_bench_serial.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", (d.get("e2e_pipeline_call") or {}).get("ms_per_call"), "parity", d.get("parity"))
        print(d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

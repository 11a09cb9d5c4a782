#!/bin/bash
# round 2, call D: whole GPU suite after the front-end fusions (K-folded conv1, stream-form statistics / planes straight from the
# raw convolution, per-hop waveform statistics); bench for both embedding variants
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $out/r2d_tests.log; tail -12 $out/r2d_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2d_bench.json 2> $out/r2d_bench.err; tail -3 $out/r2d_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --embedding wespeaker --no-cpu-baseline > $out/r2d_bench_wespeaker.json 2> $out/r2d_bench_wespeaker.err; tail -3 $out/r2d_bench_wespeaker.err
python - <<'PY'
import json
for f in ("r2d_bench.json", "r2d_bench_wespeaker.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", (d.get("e2e_pipeline_call") or {}).get("ms_per_call"), "parity", d.get("parity"))
        print(d["kernels_ms_per_step"])
        print(d["roofline"].get("step"), d["roofline"].get("largest_gemm"))
    except Exception as e: print(f, "bad", e)
PY

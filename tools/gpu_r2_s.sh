#!/bin/bash
# round 2, call S: W-resident SincNet convolutions, parallel instnorm_finalize, smaller skip grids; full suite, repeatability
# stress (fixed harness), bench with A/B (DG_NO_WRES)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $out/r2s_tests.log; tail -3 $out/r2s_tests.log
timeout 900 python tests/manual/stress_parity.py 40 256 > $out/r2s_stress.log 2>&1; echo "stress rc=$?"; tail -6 $out/r2s_stress.log | cut -c1-300
export DG_BENCH_PARITY_DETAIL=1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2s_bench.json 2> $out/r2s_bench.err; echo "bench rc=$?"; grep -E "parity_failed" $out/r2s_bench.err | cut -c1-600
DG_NO_WRES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2s_bench_nowres.json 2> $out/r2s_bench_nowres.err
DG_NO_OVERLAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2s_bench_alone.json 2> $out/r2s_bench_alone.err
python - <<'PY'
import json
for f in ["r2s_bench", "r2s_bench_nowres", "r2s_bench_alone"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        c=d.get("e2e_pipeline_call") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), "parity", d.get("parity"), d.get("clocks"))
        print("  rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

#!/bin/bash
# round 2, call V: use guards on shared model handles (new test), full suite, smoke, bench with default arguments (the driver's call)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 > $out/r2v_tests.log; tail -4 $out/r2v_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $out/r2v_bench_default.json 2> $out/r2v_bench_default.err; echo "bench rc=$?"; grep -E "parity_failed" $out/r2v_bench_default.err | cut -c1-600
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/r2v_bench_reference.json 2> $out/r2v_bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v_bench_default.json").read().strip().splitlines()[-1])
c=d.get("e2e_pipeline_call") or {}
print({k:d.get(k) for k in ("value","ms_per_step","steps","warmup")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), "parity", d.get("parity"), d.get("clocks"), "launches", d.get("gpu_launches"))
print("cpu", d.get("cpu_baseline"))
r=json.loads(open("gpurun_out/r2v_bench_reference.json").read().strip().splitlines()[-1])
print("reference arm", {k:r.get(k) for k in ("impl","value","ms_per_step","cpu_baseline")})
PY

#!/bin/bash
# round 2, call R: memcheck of one small step; then the driver's own sequence (full pytest -> smoke -> bench), bench x3 with the
# parity diagnosis, to reproduce the failure seen after the test suite in calls N and P
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 420 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 python tests/manual/one_step.py 8 > $out/r2r_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|Error" $out/r2r_memcheck.log | head -10
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $out/r2r_tests.log; tail -3 $out/r2r_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
export DG_BENCH_PARITY_DETAIL=1
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2r_bench$i.json 2> $out/r2r_bench$i.err; echo "== bench $i rc=$?"
  grep -E "parity detail|parity_failed" $out/r2r_bench$i.err | cut -c1-600 | head -5
done
python - <<'PY'
import json
for i in (1,2,3):
    try:
        d=json.loads(open(f"gpurun_out/r2r_bench{i}.json").read().strip().splitlines()[-1])
        c=d.get("e2e_pipeline_call") or {}
        print(i, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), c.get("h2d_bytes_per_step"), c.get("phases_ms_per_call"), "parity", d.get("parity"), d.get("clocks"))
        print("  rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(i, "bad", e)
PY

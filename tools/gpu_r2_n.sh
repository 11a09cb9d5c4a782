#!/bin/bash
# round 2, call N: item-aligned pooling tiles (batch invariance), cdist metrics, sub-batch timeline of __call__, ncu --set full of one
# serial step (every kernel) + summary, variant-B bench line
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
NCU=/usr/local/cuda/bin/ncu
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/r2n_tests.log; tail -6 $out/r2n_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2n_bench.json 2> $out/r2n_bench.err; tail -2 $out/r2n_bench.err
for plan in 128,128 64,128,64 256; do
  DG_CALL_TIMING=1 DG_CALL_PLAN=$plan timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-stream-leg --no-parity-check > $out/r2n_plan_$plan.json 2> $out/r2n_plan_$plan.err
  echo "plan $plan:"; grep -A3 "dg_pipeline_call_host\|sub-batch" $out/r2n_plan_$plan.err | tail -5
done
DG_TRACE_LAUNCHES=1 DG_NO_OVERLAP=1 timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/r2n_step python tests/manual/one_step.py > $out/r2n_ncu.log 2> $out/r2n_trace.log
$NCU -i /tmp/r2n_step.ncu-rep --page raw --csv > $out/r2n_step_raw.csv 2>/dev/null
python tools/ncu_summary.py $out/r2n_step_raw.csv $out/r2n_trace.log $out/r2n_step > $out/r2n_step_table.log 2>&1; head -40 $out/r2n_step_table.log
timeout 600 python bench.py --steps 10 --warmup 3 --embedding wespeaker --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2n_bench_wespeaker.json 2> $out/r2n_bench_wespeaker.err
python - <<'PY'
import json
for f in ["gpurun_out/r2n_bench.json", "gpurun_out/r2n_bench_wespeaker.json", "gpurun_out/r2n_plan_128,128.json", "gpurun_out/r2n_plan_64,128,64.json", "gpurun_out/r2n_plan_256.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get("e2e_pipeline_call") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), c.get("phases_ms_per_call"), "parity", d.get("parity"))
        if "plan" not in f: print("  rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

#!/bin/bash
# round 2, call M: conv1 / conv2 with MaxPool3 + InstanceNorm partial sums in the GEMM epilogue, power-of-two weight-plane scale,
# LSTM shape by batch size, persistent gather pool + sub-batch plans of dg_pipeline_call_host
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $out/r2m_tests.log; tail -6 $out/r2m_tests.log
DG_CALL_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2m_bench.json 2> $out/r2m_bench.err; grep dg_pipeline_call_host $out/r2m_bench.err | tail -2
DG_NO_POOL3_FUSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2m_bench_nopool3.json 2> $out/r2m_bench_nopool3.err
DG_NO_WSCALE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2m_bench_nowscale.json 2> $out/r2m_bench_nowscale.err
for plan in 128,128 88,88,80 64,96,96 96,96,64 64,128,64 48,104,104; do
  DG_CALL_TIMING=1 DG_CALL_PLAN=$plan timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-stream-leg --no-parity-check > $out/r2m_plan_$plan.json 2> $out/r2m_plan_$plan.err
  echo "plan $plan: $(grep dg_pipeline_call_host $out/r2m_plan_$plan.err | tail -1)"
done
python - <<'PY'
import json, glob
for f in ["gpurun_out/r2m_bench.json", "gpurun_out/r2m_bench_nopool3.json", "gpurun_out/r2m_bench_nowscale.json"] + sorted(glob.glob("gpurun_out/r2m_plan_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get("e2e_pipeline_call") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), c.get("phases_ms_per_call"), "parity", d.get("parity"))
        if "plan" not in f: print("  rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

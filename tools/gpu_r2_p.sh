#!/bin/bash
# round 2, call P: repeatability stress (the parity failure seen once in call N), then tests / bench / ncu of the real path
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
NCU=/usr/local/cuda/bin/ncu
timeout 600 python tests/manual/stress_parity.py 40 256 > $out/r2p_stress.log 2>&1; echo "stress rc=$?"; tail -12 $out/r2p_stress.log
if ! grep -q " 0 mismatching" $out/r2p_stress.log; then
  for v in DG_NO_POOL3_FUSE=1 DG_LSTM_ROWS=16 STRESS_NO_CALLS=1 DG_CALL_NO_DEDUP=1 DG_STREAM_SINC=0 DG_GEMM_1CTA=1; do
    timeout 600 env $v python tests/manual/stress_parity.py 40 256 > $out/r2p_stress_$v.log 2>&1; echo "stress $v rc=$?"; tail -4 $out/r2p_stress_$v.log
  done
fi
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > $out/r2p_tests.log; tail -6 $out/r2p_tests.log
DG_CALL_TIMING=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2p_bench.json 2> $out/r2p_bench.err; grep -A3 "dg_pipeline_call_host" $out/r2p_bench.err | tail -5; tail -2 $out/r2p_bench.err
DG_TRACE_LAUNCHES=1 timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/r2p_step python tests/manual/one_step.py > $out/r2p_ncu.log 2> $out/r2p_trace.log
$NCU -i /tmp/r2p_step.ncu-rep --page raw --csv > /tmp/r2p_step_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r2p_step_raw.csv $out/r2p_trace.log $out/r2p_step > $out/r2p_step_table.log 2>&1; head -45 $out/r2p_step_table.log
python - <<'PY'
import json
for f in ["gpurun_out/r2p_bench.json"]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        c=d.get("e2e_pipeline_call") or {}
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), c.get("h2d_bytes_per_step"), c.get("phases_ms_per_call"), "parity", d.get("parity"))
        print("  rec", d["roofline"]["recurrence"]["us_per_dependent_step"], d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

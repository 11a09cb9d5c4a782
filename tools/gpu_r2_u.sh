#!/bin/bash
# round 2, call U: full suite (VAD on the device post-path), ncu --set full of one step of the final kernels + summary, bench
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
NCU=/usr/local/cuda/bin/ncu
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 > $out/r2u_tests.log; tail -4 $out/r2u_tests.log
DG_TRACE_LAUNCHES=1 timeout 1200 $NCU --set full --clock-control none --import-source on --profile-from-start off -f -o /tmp/r2u_step python tests/manual/one_step.py > $out/r2u_ncu.log 2> $out/r2u_trace.log
$NCU -i /tmp/r2u_step.ncu-rep --page raw --csv > /tmp/r2u_step_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r2u_step_raw.csv $out/r2u_trace.log $out/r2u_step > $out/r2u_step_table.log 2>&1; head -12 $out/r2u_step_table.log | cut -c1-220
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2u_bench.json 2> $out/r2u_bench.err; echo "bench rc=$?"; grep -E "parity_failed" $out/r2u_bench.err | cut -c1-600
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2u_bench.json").read().strip().splitlines()[-1])
c=d.get("e2e_pipeline_call") or {}
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", c.get("ms_per_call"), "parity", d.get("parity"), d.get("clocks"))
print("cpu", d.get("cpu_baseline"))
print("roofline", {k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk in ("kernel","frac","achieved","us_per_dependent_step","traffic")}) for k,v in d["roofline"].items() if k!="note"})
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final2_smoke.log
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final2_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/final2_bench.json 2> gpurun_out/final2_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final2_ref.json 2>> gpurun_out/final2_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final2_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"))
print(json.dumps(d["roofline"])[:900])
PY
tail -c 300 gpurun_out/final2_ref.json

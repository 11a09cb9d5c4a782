#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final3_smoke.log
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2 | tee gpurun_out/final3_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/final3_bench.json 2> gpurun_out/final3_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final3_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"], "cpu", d.get("cpu_baseline",{}).get("value"), d["roofline"]["kernel"], d["roofline"]["frac"])
PY

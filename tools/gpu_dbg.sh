#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed\|Error" | tail -12 | tee gpurun_out/dbg_pipe.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1g_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], d["clocks"])
PY

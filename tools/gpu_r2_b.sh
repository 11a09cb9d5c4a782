#!/bin/bash
# round 2, call B: device post-path, ring-buffer stream, drop-in pipeline tests; bench with the new legs
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_post.py tests/test_gpu_stream.py tests/test_gpu_pipeline.py tests/test_gpu_batch256.py -q -m gpu -x 2>&1 | tail -40 > $out/r2b_tests.log; tail -12 $out/r2b_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2b_bench.json 2> $out/r2b_bench.err; tail -3 $out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", d.get("e2e_stream"), "call", d.get("e2e_pipeline_call"), "parity", d.get("parity"))
    print(d["kernels_ms_per_step"])
except Exception as e: print("bad", e)
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== default" | tee gpurun_out/stream_tests.log
timeout 300 python -m pytest tests/test_gpu_nets.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -2 | tee -a gpurun_out/stream_tests.log
echo "== DG_STREAM_SINC=1" | tee -a gpurun_out/stream_tests.log
DG_STREAM_SINC=1 timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_shapes.py -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed\|Error" | tail -10 | tee -a gpurun_out/stream_tests.log
DG_STREAM_SINC=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/stream_bench.json 2> gpurun_out/stream_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/stream_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], json.dumps(d["kernels_ms_per_step"]))
PY

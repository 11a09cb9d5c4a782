#!/bin/bash
# round 2, call Z: clustering kernel with float64 embeddings converted once per chunk
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_shared_identity.py tests/test_gpu_batch256.py tests/test_gpu_pipeline.py -q -m gpu 2>&1 | tail -3
DG_CLUSTER_TIMING=1 timeout 120 python tests/manual/cluster_timing.py 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2z_bench.json 2> $out/r2z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "parity", d.get("parity"))
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["kernels_ms_per_step"])
PY

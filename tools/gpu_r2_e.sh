#!/bin/bash
# round 2, call E: fused TDNN5 + statistics pooling, sub-batched __call__; tests + bench + A/B
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_nets.py tests/test_gpu_pipeline.py tests/test_gpu_batch256.py tests/test_gpu_stream.py tests/test_gpu_post.py tests/test_gpu_shapes.py tests/test_shared_identity.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 > $out/r2e_tests.log; tail -25 $out/r2e_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/r2e_bench.json 2> $out/r2e_bench.err; tail -3 $out/r2e_bench.err
DG_NO_POOL_FUSE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2e_bench_nofuse.json 2> $out/r2e_bench_nofuse.err
for sp in 1 3; do DG_CALL_SPLIT=$sp timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stream-leg --no-parity-check > $out/r2e_bench_split$sp.json 2> $out/r2e_bench_split$sp.err; done
python - <<'PY'
import json
for f in ("r2e_bench.json", "r2e_bench_nofuse.json", "r2e_bench_split1.json", "r2e_bench_split3.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", (d.get("e2e_pipeline_call") or {}).get("ms_per_call"), "parity", d.get("parity"))
        print(d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

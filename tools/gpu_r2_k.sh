#!/bin/bash
# round 2, call K: CTA-pair GEMM (cta_group::2) bring-up + TC_POOL trims
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -q -m gpu -x 2>&1 | tail -15 > $out/r2k_gemm.log; tail -8 $out/r2k_gemm.log
timeout 1200 python -m pytest tests/test_gpu_nets.py tests/test_gpu_pipeline.py tests/test_gpu_shapes.py tests/test_gpu_batch256.py tests/test_zz_wespeaker.py -q -m gpu 2>&1 | tail -15 > $out/r2k_tests.log; tail -8 $out/r2k_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $out/r2k_bench.json 2> $out/r2k_bench.err; tail -3 $out/r2k_bench.err
DG_GEMM_1CTA=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-stream-leg --no-pipeline-call > $out/r2k_bench_1cta.json 2> $out/r2k_bench_1cta.err
DG_NO_OVERLAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2k_bench_alone.json 2> $out/r2k_bench_alone.err
python - <<'PY'
import json
for f in ("r2k_bench.json", "r2k_bench_1cta.json", "r2k_bench_alone.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"), "call", (d.get("e2e_pipeline_call") or {}).get("ms_per_call"), "parity", d.get("parity"))
        print(d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY

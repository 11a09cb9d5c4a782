#!/bin/bash
# A/B run on the GPU box: parity tests of the default path, then bench lines for the kernel variants.
# usage: tools/gpu_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out
cd "$(dirname "$0")/.."
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_nets.py tests/test_gpu_pipeline.py tests/test_gpu_shapes.py -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed\|shape (" | tail -40 | tee $out/${tag}_tests.log
b() { name=$1; shift; timeout 240 python bench.py --no-cpu-baseline "$@" > $out/${tag}_bench_$name.json 2>> $out/${tag}_bench.err; }
b default --steps 10 --warmup 3
DG_NO_V8=1 b nov8 --steps 10 --warmup 3
b serial --steps 20 --warmup 3 --serial
DG_NO_V8=1 b serial_nov8 --steps 20 --warmup 3 --serial
b b1 --steps 30 --warmup 5 --serial --batch 1
for f in $out/${tag}_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d.get("e2e",{}).get("value"), json.dumps(d.get("kernels_ms_per_step")))
except Exception as e: print("bad", e)
PY
done

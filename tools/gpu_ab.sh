#!/bin/bash
# A/B run on the GPU box: parity tests of the default path, then bench lines for the kernel variants.
# usage: tools/gpu_ab.sh <tag>
tag=${1:-ab}
out=gpurun_out
mkdir -p $out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_nets.py tests/test_gpu_pipeline.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 > $out/${tag}_tests.log
tail -5 $out/${tag}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
DG_LSTM_V2=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_lstm2.json 2>> $out/${tag}_bench.err
DG_SPLIT_BF16=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_bf16.json 2>> $out/${tag}_bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --serial > $out/${tag}_bench_serial.json 2>> $out/${tag}_bench.err
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --serial --batch 1 > $out/${tag}_bench_b1.json 2>> $out/${tag}_bench.err
DG_SPLIT_BF16=1 DG_LSTM_V2=1 timeout 600 python -m pytest tests/test_gpu_nets.py -x -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed" | tail -8 > $out/${tag}_tests_old.log
for f in $out/${tag}_bench*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d.get("e2e",{}).get("value"), json.dumps(d.get("kernels_ms_per_step")))
except Exception as e: print("bad", e)
PY
done

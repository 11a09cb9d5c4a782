#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_shapes.py -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed" | tail -12 > gpurun_out/dbg_nets.log
DG_LSTM_TIMING=1 timeout 300 python bench.py --serial --steps 2 --warmup 1 --no-cpu-baseline 2>&1 >/dev/null | grep "lstm_rec timing" | head -3 > gpurun_out/lstm_timing.log
DG_LSTM_TIMING=1 timeout 300 python bench.py --serial --steps 2 --warmup 1 --no-cpu-baseline --batch 1 2>&1 >/dev/null | grep "lstm_rec timing" | head -3 >> gpurun_out/lstm_timing.log
cat gpurun_out/dbg_nets.log gpurun_out/lstm_timing.log
b() { name=$1; shift; timeout 240 python bench.py --no-cpu-baseline "$@" > gpurun_out/r1f_bench_$name.json 2>> gpurun_out/r1f_bench.err; }
b default --steps 10 --warmup 3
b serial --steps 20 --warmup 3 --serial
b b1 --steps 30 --warmup 5 --serial --batch 1
for f in gpurun_out/r1f_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d.get("e2e",{}).get("value"), "lstm_rec", d.get("kernels_ms_per_step",{}).get("lstm_rec"))
except Exception as e: print("bad", e)
PY
done

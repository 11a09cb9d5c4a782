#!/bin/bash
# round 2, call I: kernel-alone timings (DG_NO_OVERLAP=1: one stream, back to back) and an ncu launch list
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
DG_NO_OVERLAP=1 timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2i_bench_alone.json 2> $out/r2i_bench_alone.err
DG_NO_OVERLAP=1 DG_NO_POOL_FUSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2i_bench_alone_nofuse.json 2> $out/r2i_bench_alone_nofuse.err
NCU=/usr/local/cuda/bin/ncu
DG_NO_OVERLAP=1 timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $out/r2i_launches.csv python bench.py --steps 1 --warmup 1 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check > $out/r2i_ncu.log 2>&1
python - <<'PY'
import json
for f in ("r2i_bench_alone.json", "r2i_bench_alone_nofuse.json"):
    try:
        d=json.loads(open('gpurun_out/'+f).read().strip().splitlines()[-1])
        print(f, {k:d.get(k) for k in ("value","ms_per_step")})
        print(d["kernels_ms_per_step"])
    except Exception as e: print(f, "bad", e)
PY
tail -5 $out/r2i_launches.csv | cut -c1-300

#!/bin/bash
# round 2: two GPUs, one stream per GPU (config #4) + the shared-identity leg (config #5) in the same line
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $out/r2_bench_n$N.json 2> $out/r2_bench_n$N.err; echo "rc=$?"; tail -3 $out/r2_bench_n$N.err
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2_bench_n{n}.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "stream", (d.get("e2e_stream") or {}).get("value"))
print("shared_identity", d.get("shared_identity"))
print("parity", d.get("parity"), d.get("clocks"))
PY

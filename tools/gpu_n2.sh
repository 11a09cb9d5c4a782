#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err; tail -c 600 gpurun_out/n2_$name.json | head -c 600; echo; }
run bench --steps 10 --warmup 3
run shared --steps 10 --warmup 3 --shared-identity
timeout 300 python -m pytest tests/test_shared_identity.py -q -m gpu 2>&1 | tail -2
python - <<'PY'
import json
for n in ("bench","shared"):
    try:
        d=json.loads(open(f'gpurun_out/n2_{n}.json').read().strip().splitlines()[-1]); print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["n_gpus"])
    except Exception as e: print(n,"bad",e)
PY

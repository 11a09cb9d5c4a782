#!/bin/bash
# Round-end validation on the GPU box: smoke, the full GPU suite, the bench lines, the ncu launch list and
# --set full captures of the top kernels.  Everything lands in gpurun_out/final_*.
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/final_smoke.log 2>&1; tail -1 $out/final_smoke.log
timeout 1200 python -m pytest tests -q -m gpu -s 2>&1 | grep -i "err\|passed\|failed\|error" | tail -30 > $out/final_tests.log; tail -3 $out/final_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > $out/final_bench.json 2> $out/final_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $out/final_ref.json 2>> $out/final_bench.err
timeout 200 python bench.py --steps 20 --warmup 3 --serial --no-cpu-baseline > $out/final_bench_serial.json 2>> $out/final_bench.err
timeout 200 python bench.py --steps 30 --warmup 5 --serial --batch 1 --no-cpu-baseline > $out/final_bench_b1.json 2>> $out/final_bench.err
timeout 200 python bench.py --steps 30 --warmup 5 --serial --batch 32 --no-cpu-baseline > $out/final_bench_b32.json 2>> $out/final_bench.err
NCU=/usr/local/cuda/bin/ncu
B="python bench.py --steps 2 --warmup 1 --serial --no-cpu-baseline"
timeout 400 $NCU --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/final_launches.csv $B > $out/ncu_b.log 2>&1
full() { name=$1; pat=$2; skip=$3; timeout 300 $NCU --set full --clock-control none --import-source on -k "regex:$pat" -s $skip -c 1 -f -o $out/final_$name $B > $out/ncu_$name.log 2>&1; }
full lstm 'lstm_tc3' 5
full sinc0 'sinc0_tc' 2
# (-k matches the base name only: the templated GEMM launches are picked by index)
timeout 400 $NCU --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 8 -f -o $out/final_gemm $B > $out/ncu_gemm.log 2>&1
ls -la $out/final_* | awk '{print $5, $9}'
for f in $out/final_bench*.json $out/final_ref.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step")}, "e2e", d.get("e2e",{}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "clocks", d.get("clocks"))
except Exception as e: print("bad", e)
PY
done

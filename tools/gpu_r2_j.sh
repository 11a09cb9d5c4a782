#!/bin/bash
# round 2, call J: ncu --set full of the fused TDNN5 + pooling GEMM (and the un-fused one for comparison)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
NCU=/usr/local/cuda/bin/ncu
B="python bench.py --steps 1 --warmup 1 --serial --no-cpu-baseline --no-stream-leg --no-pipeline-call --no-parity-check"
DG_NO_OVERLAP=1 timeout 600 $NCU --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 24 -c 12 -f -o $out/r2j_gemm $B > $out/r2j_ncu.log 2>&1
$NCU -i $out/r2j_gemm.ncu-rep --page raw --csv > $out/r2j_gemm_raw.csv 2>/dev/null
$NCU -i $out/r2j_gemm.ncu-rep --page source --csv --kernel-name regex:gemm_tc_kernel > $out/r2j_gemm_source.csv 2>/dev/null
ls -la $out/r2j_*; python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2j_gemm_raw.csv')))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
idx=[i for i,h in enumerate(hdr) if any(h==w or h.startswith(w) for w in want)]
for r in rows[2:]:
    print([ (hdr[i][:40], r[i][:60]) for i in idx])
PY

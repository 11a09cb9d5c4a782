#!/bin/bash
# round 2, call W: probe of the first processes after the GPU test suite (mismatches of 4e-4 .. 2e-2 seen only in that slot)
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
probe() { timeout 300 python tests/manual/after_suite_probe.py "$1" 2>&1 | grep -E "^$1|Error|error" | cut -c1-400; DG_SELFTEST_AMP=0.01 timeout 120 python - "$1" <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
from diart_b200 import _lib
lib = _lib.lib()
d, r = C.c_float(), C.c_float()
_lib.check(lib.dg_selftest_gemm_tc(4096, 256, 1, 1, 1024, 0, C.byref(d), C.byref(r)))
print(f"{sys.argv[1]}: tiny-operand gemm: max diff {d.value:.3e} rms {r.value:.3e} rel {d.value / r.value:.2e}")
PY
}
probe "cold"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
probe "after-suite-1"
probe "after-suite-2"
timeout 600 python -m pytest tests/test_zz_wespeaker.py -q -m gpu 2>&1 | tail -2
probe "after-wespeaker"
timeout 600 python -m pytest tests/test_gpu_batch256.py tests/test_gpu_nets.py -q -m gpu 2>&1 | tail -2
probe "after-nets"
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
probe "after-suite-smoke"

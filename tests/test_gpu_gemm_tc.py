"""The tcgen05/TMEM/TMA shifted-window GEMM (bf16x3 split precision) against the float32 SIMT GEMM on the
same seeded random operands, layer shapes of the path.  Bar: max |diff| < 3e-5 x output RMS x sqrt(K/64)
(three bf16 products carry 16 significand bits per operand)."""
import ctypes as C

import pytest

from diart_b200 import _lib

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, Cin, KW, dil, N, epi)
    (1000, 64, 1, 1, 256, 0),        # one k-block, ragged M
    (2368, 64, 5, 1, 512, 1),        # tdnn1-like, split epilogue
    (1184, 512, 3, 2, 512, 1),       # tdnn2 (dilation 2)
    (1184, 512, 3, 3, 512, 2),       # tdnn3 (dilation 3), f32 epilogue
    (1184, 512, 1, 1, 1500, 2),      # tdnn5: N not a multiple of the tile
    (4096, 256, 1, 1, 1024, 0),      # LSTM input projection
    (300, 128, 2, 7, 128, 0),        # BN=128 variant, odd dilation
    (5328, 128, 5, 1, 64, 0),        # SincNet conv1 (80 -> 128 padded channels, N = 64 of a 128 tile)
    (768, 3008, 1, 1, 512, 0),       # Linear(3000, 512) with K padded to 47 k-blocks
]


@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_tc_matches_simt(shape, cuda_device):
    M, Cin, KW, dil, N, epi = shape
    diff, rms = C.c_float(), C.c_float()
    _lib.check(_lib.lib().dg_selftest_gemm_tc(M, Cin, KW, dil, N, epi, C.byref(diff), C.byref(rms)))
    tol = 3e-5 * rms.value * (KW * Cin / 64) ** 0.5
    print(f"shape {shape}: max abs diff {diff.value:.3e}, output rms {rms.value:.3e}, tol {tol:.3e}")
    assert diff.value == diff.value, "NaN in the comparison"
    assert diff.value < tol

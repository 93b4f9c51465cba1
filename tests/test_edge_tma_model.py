"""CPU model of the edge-window patch of tc_conv's experimental tensor-copy path (TVC_TC_EDGE_TMA=1, csrc/tc_conv.cu
`ETMA`): a window fetched as a plain shifted view of the plane (neighbouring utterances' rows, zeros outside the tensor)
and patched with the kernel's two formulas must equal the clamped (replicate-padded) gather it replaces, slot by slot."""
import numpy as np
import pytest


@pytest.mark.parametrize("T,dil", [(108, 1), (108, 27), (432, 9), (130, 3), (96 + 1, 27), (1728, 27), (127, 64)])
def test_patched_window_equals_clamped_gather(T, dil):
    B, tile = 3, 128
    rows = B * T
    x = np.arange(1, rows + 1, dtype=np.float64)            # every operand row distinct and non-zero
    win = tile + 2 * dil
    tpu = -(-T // tile)
    for b in range(B):
        baseT = b * T
        for k in range(tpu):
            tt0 = k * tile
            org = tt0 - dil
            # what the producers gather today
            want = np.array([x[baseT + min(max(org + m, 0), T - 1)] for m in range(win)])
            # tensor copy: shifted view of the whole plane, zeros outside it
            got = np.array([x[baseT + org + m] if 0 <= baseT + org + m < rows else 0.0 for m in range(win)])
            # the MMA warp's patch (tc_conv.cu, `if constexpr (ETMA)`)
            pad_lo = -org if org < 0 else 0
            hi_first = T - org
            if pad_lo > 0:
                got[:pad_lo] = got[pad_lo]
            if hi_first < win:
                got[hi_first:] = got[hi_first - 1]
            assert np.array_equal(got, want), (b, k)

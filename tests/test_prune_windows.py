"""The decoder's output-pruning plan (tvc_decoder_plan_windows = nets_tc.cu decoder_plan_windows, host arithmetic of the shipped
library) against a brute-force dependency trace, on the CPU.

For a kept sample range [t0, t1) the fused full-rate block walks the 426-sample windows that produce kept samples; every window
row resamples two rows of level 3's output (F.interpolate x5: ATen's fp32 coordinate arithmetic).  A pruned level
i works on rows [wa, wb) of its utterances as if they were the whole utterance, so a row of its output is exact only if it is
at least 40 rows (the reach of its k = 3 convs with dilations 1, 3, 9, 27) away from a cut edge; its resampler reads two rows of
level i - 1 per output row.  The test traces what is really read, level by level, and checks that everything read lies in the
exact part of the producing level's window."""
import ctypes

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

FAC = [2, 3, 4, 4, 5]            # kUpFac: L/480 -> L/240 -> L/80 -> L/20 -> L/5 -> L
f32 = np.float32


def lin_rows(t, scale, in_len):
    """Rows i0, i1 that F.interpolate(mode='linear') reads for output positions t (fp32, as tvc_common.cuh lin_coord)."""
    src = (f32(scale) * (t.astype(f32) + f32(0.5))).astype(f32) - f32(0.5)       # fma vs mul+add: differs in the last bit only;
    src = np.maximum(src, f32(0))                                               # both candidates are traced below
    i0 = np.minimum(np.floor(src).astype(np.int64), in_len - 1)
    i1 = i0 + (i0 < in_len - 1)
    return i0, i1


def plan(Lf, t0, t1):
    from tinyvc_b200 import _lib
    wa, wb = (ctypes.c_int32 * 5)(), (ctypes.c_int32 * 5)()
    _lib.check(_lib.lib().tvc_decoder_plan_windows(Lf, t0, t1, wa, wb), "tvc_decoder_plan_windows")
    return list(wa), list(wb)


def check(Lf, t0, t1):
    L = Lf * 480
    T = [Lf * 2, Lf * 6, Lf * 24, Lf * 96, L]
    wa, wb = plan(Lf, t0, t1)
    assert wa[4] == 0 and wb[4] == L
    for i in range(4):
        assert 0 <= wa[i] < wb[i] <= T[i]
    # rows of level 3 the walked windows of the block read (window k: block rows k*426-44 .. k*426-43+512, clamped)
    k_lo, k_hi = t0 // 426, (t1 - 1) // 426
    rows = np.unique(np.clip(np.concatenate([np.arange(k * 426 - 44, k * 426 - 43 + 513) for k in range(k_lo, k_hi + 1)]), 0, L - 1))
    need = np.unique(np.concatenate(lin_rows(rows, 1.0 / 5.0, T[3])))
    for i in (3, 2, 1, 0):
        full = wa[i] == 0 and wb[i] == T[i]
        if full:
            break                                   # everything below a full level is full as well
        assert wb[i] - wa[i] >= 384                 # pruned levels keep the shared-window tiling of the full run
        # output row r of this level is exact iff the rows r - 40 .. r + 40 of its resampled input (clipped to the utterance:
        # its true ends carry the true replicate padding) are all inside the window and exact themselves
        ext = np.arange(max(0, int(need.min()) - 40), min(T[i], int(need.max()) + 41))
        assert ext.min() >= wa[i] and ext.max() < wb[i], (Lf, t0, t1, i, wa, wb, int(need.min()), int(need.max()))
        # the resampler reads two rows of level i - 1 (level 0: of the frame-rate tensor, always complete) per row: those read
        # for `ext` must be exact below; those read for the rest of the window must merely exist in the compact tensor
        in_len = T[i - 1] if i > 0 else Lf
        if i > 0 and not (wa[i - 1] == 0 and wb[i - 1] == T[i - 1]):
            touched = np.concatenate(lin_rows(np.arange(wa[i], wb[i]), 1.0 / FAC[i], in_len))
            assert touched.min() >= wa[i - 1] and touched.max() < wb[i - 1]
        need = np.unique(np.concatenate(lin_rows(ext, 1.0 / FAC[i], in_len)))


def test_streaming_tick_windows():
    """The tick (28 frames, keeps y[-9600:-3840]): levels 3 and 2 are pruned to about half / 60 %, the rest is full."""
    wa, wb = plan(28, 3840, 9600)
    assert (wa[3], wb[3]) == (717, 2010) and (wa[2], wb[2]) == (138, 544)
    assert (wa[1], wb[1], wa[0], wb[0]) == (0, 168, 0, 56)
    check(28, 3840, 9600)


def test_full_range_is_never_pruned():
    for Lf in (1, 18, 500):
        wa, wb = plan(Lf, 0, Lf * 480)
        assert wa == [0] * 5 and wb == [Lf * 2, Lf * 6, Lf * 24, Lf * 96, Lf * 480]


@settings(max_examples=300, deadline=None, derandomize=True)
@given(Lf=st.integers(1, 700), a=st.floats(0, 1), b=st.floats(0, 1), short=st.booleans())
def test_windows_cover_every_dependency(Lf, a, b, short):
    L = Lf * 480
    t0 = min(int(a * L), L - 1)
    t1 = min(L, t0 + 1 + int(b * (4000 if short else L)))
    check(Lf, t0, t1)

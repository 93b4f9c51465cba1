"""GPU parity of the tcgen05 tensor-core Conv1d kernel (csrc/tc_conv.cu) through the C-ABI probe.

Oracle: torch-CPU float64 `F.conv1d` on replicate-padded input, exactly how nn.Conv1d with
padding_mode='replicate' evaluates the decoder's convs (module/tinyvc/decoder.py:143-146,165-171),
plus FiLM (decoder.py:88-97) and the Downsample residual 1x1 (decoder.py:143,157).
Tolerance: the kernel evaluates x*w as three bf16 products with fp32 accumulation (relative error
~2^-16 per term); relative RMS error must stay below 3e-5 (TF32 would be ~3e-4).
"""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(x, w, b, dil, aux_x, aux_w, aux_b, aux_mode, res, epi_act):
    xd, wd = x.double(), w.double()
    k = w.shape[2]
    pad = dil * (k - 1) // 2
    y = F.conv1d(F.pad(xd, (pad, pad), mode="replicate") if pad else xd, wd, b.double(), dilation=dil)
    cout = w.shape[0]
    if aux_mode == 1:
        y = y + F.conv1d(aux_x.double(), aux_w.double()[:, :, None], aux_b.double())
    elif aux_mode == 2:
        ss = F.conv1d(aux_x.double(), aux_w.double().reshape(2 * cout, -1)[:, :, None], aux_b.double().reshape(-1))
        y = y * ss[:, :cout] + ss[:, cout:]
    if res is not None:
        y = y + res.double()
    if epi_act == 2:
        y = F.gelu(y)
    elif epi_act == 3:
        y = F.elu(y) + 1
    return y


# Padded mode (tc_conv.cuh): the input planes carry `pin` stored replicate rows per utterance side, the plane output
# `pout`; the probe also checks that the epilogue wrote every replicate row of the output.  (case, (pin, pout))
PAD_CASES = [
    ((4, 36, 384, 384, 3, 1, 0, 0, False, 0, 1, 48), (1, 3)),
    ((4, 36, 384, 384, 3, 3, 384, 2, True, 0, 1, 48), (3, 9)),
    ((4, 36, 384, 384, 3, 27, 384, 2, True, 0, 0, 48), (27, 0)),
    ((3, 108, 192, 192, 3, 9, 0, 0, False, 0, 1, 64), (9, 27)),
    ((5, 2, 384, 384, 3, 27, 384, 2, True, 0, 0, 48), (27, 0)),
    ((7, 1, 48, 48, 3, 2, 0, 0, False, 0, 1, 48), (2, 4)),
    ((3, 100, 48, 96, 3, 4, 48, 1, False, 0, 0, 48), (4, 0)),
    ((2, 432, 96, 96, 3, 9, 0, 0, False, 0, 1, 48), (9, 27)),
    ((2, 1000, 24, 24, 3, 27, 24, 2, True, 0, 1, 24), (27, 5)),
    ((9, 33, 17, 24, 3, 1, 0, 0, False, 0, 0, 24), (6, 1)),
]

CASES = [
    # (B, T, Cin, Cout, K, dil, aux_cin, aux_mode, use_res, epi_act, out_act, NT)
    (2, 300, 24, 24, 3, 1, 0, 0, False, 0, 1, 24),
    (2, 300, 24, 24, 3, 27, 24, 2, True, 0, 1, 24),
    (3, 100, 48, 96, 3, 4, 48, 1, False, 0, 0, 96),
    (8, 18, 768, 512, 1, 1, 0, 0, False, 0, 0, 128),
    (4, 36, 384, 384, 3, 9, 384, 2, True, 0, 1, 64),
    (5, 18, 128, 976, 1, 1, 0, 0, False, 3, 0, 128),
    (2, 777, 17, 24, 3, 1, 0, 0, False, 0, 0, 24),
    (3, 18, 961, 961, 1, 1, 0, 0, False, 0, 0, 128),
    (1, 5, 24, 24, 3, 3, 0, 0, False, 0, 0, 24),
    (2, 129, 192, 96, 1, 1, 0, 0, True, 2, 0, 96),
]


@pytest.mark.parametrize("case,pads", PAD_CASES, ids=[f"p{i}" for i in range(len(PAD_CASES))])
def test_tc_conv_padded_mode(case, pads):
    _check(case, pads)


@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_tc_conv_matches_fp64(case):
    _check(case)


def _check(case, pads=None):
    from tinyvc_b200 import _lib
    B, T, Cin, Cout, K, dil, aux_cin, aux_mode, use_res, epi_act, out_act, NT = case
    g = torch.Generator().manual_seed(hash(case) % (2**31))
    x = torch.randn(B, Cin, T, generator=g)
    w = torch.randn(Cout, Cin, K, generator=g) / (Cin * K) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    aux_x = aux_w = aux_b = None
    if aux_mode == 1:
        aux_x = torch.randn(B, aux_cin, T, generator=g)
        aux_w = torch.randn(Cout, aux_cin, generator=g) / aux_cin**0.5
        aux_b = torch.randn(Cout, generator=g) * 0.1
    elif aux_mode == 2:
        aux_x = torch.randn(B, aux_cin, T, generator=g)
        aux_w = torch.randn(2, Cout, aux_cin, generator=g) / aux_cin**0.5
        aux_b = torch.randn(2, Cout, generator=g) * 0.1
    res = torch.randn(B, Cout, T, generator=g) if use_res else None
    ref = _ref(x, w, b, dil, aux_x, aux_w, aux_b, aux_mode, res, epi_act)

    dev = torch.device("cuda")
    xd = x.to(dev).contiguous()
    auxd = aux_x.to(dev).contiguous() if aux_x is not None else None
    resd = res.to(dev).contiguous() if res is not None else None
    y = torch.full((B, Cout, T), float("nan"), device=dev)
    yp = torch.full((B, Cout, T), float("nan"), device=dev)
    wc, bc = w.contiguous(), b.contiguous()
    awc = aux_w.contiguous() if aux_w is not None else None
    abc = aux_b.contiguous() if aux_b is not None else None
    ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None  # noqa: E731
    if pads:
        _lib.set_option("probe_pad", f"{pads[0]},{pads[1]}")
    try:
        rc = _lib.lib().tvc_tc_conv_probe(ptr(xd), ptr(wc), ptr(bc), B, T, Cin, Cout, K, dil, ptr(auxd), ptr(awc), ptr(abc),
                                          aux_cin, aux_mode, ptr(resd), epi_act, out_act, NT, ptr(y), ptr(yp), None)
    finally:
        _lib.set_option("probe_pad", "0,0")
    _lib.check(rc, "tvc_tc_conv_probe")
    torch.cuda.synchronize()
    got = y.cpu().double()
    scale = ref.pow(2).mean().sqrt()
    err = (got - ref).pow(2).mean().sqrt() / scale
    mx = (got - ref).abs().max() / scale
    print(f"[tc_conv] case={case} rel_rms={err:.3e} rel_max={mx:.3e}")
    assert torch.isfinite(got).all()
    assert err < 3e-5, (case, float(err))
    refp = F.leaky_relu(ref, 0.1) if out_act == 1 else ref
    errp = (yp.cpu().double() - refp).pow(2).mean().sqrt() / scale
    assert errp < 3e-5, (case, float(errp))

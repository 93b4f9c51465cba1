"""The three bit-level numerics contracts the CUDA kernels implement,
restated in NumPy (oracle/numerics_np.py), checked against torch-CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import numerics_np as N
from oracle import tinyvc_oracle as O
from tinyvc_b200 import synth


@pytest.mark.parametrize("tin,tout,sf", [(18, 8640, None), (200, 96000, None), (7, 3360, 480), (36, 72, 2), (72, 216, 3),
                                         (216, 864, 4), (1728, 8640, 5), (8640, 1728, 1 / 5), (1728, 432, 1 / 4),
                                         (108, 36, 1 / 3), (75, 4800, None), (1, 480, None)])
def test_interp_contract_bit_exact(tin, tout, sf):
    g = torch.Generator().manual_seed(tin * 7 + tout)
    x = torch.randn(3, 2, tin, generator=g)
    if sf is None:
        ref = F.interpolate(x, tout, mode="linear")
    else:
        ref = F.interpolate(x, scale_factor=sf, mode="linear")
        assert ref.shape[-1] == tout
    got = N.interp_linear(x.numpy(), tout, scale_factor=sf)
    assert np.array_equal(got, ref.numpy())


@pytest.mark.parametrize("lf,seed", [(18, 0), (57, 1), (200, 2)])
def test_harmonic_phase_contract_bit_exact(lf, seed):
    f0 = synth.synth_f0(1, lf, torch.Generator().manual_seed(seed))          # [1,1,Lf]
    _, theta_ref = O.oscillate_harmonics(f0, return_theta=True)
    theta, uv = N.harmonic_theta(f0[0, 0].numpy())
    assert np.array_equal(theta, theta_ref[0].numpy())
    uv_ref = F.interpolate((f0 > 20.0).float(), lf * 480, mode="linear")[0, 0].numpy()
    assert np.array_equal(uv, uv_ref)


def test_harmonic_phase_contract_high_pitch():
    """Random-init pitch estimators emit f0 up to ~14 kHz: |I| reaches 1e6."""
    g = torch.Generator().manual_seed(5)
    f0 = (torch.rand(1, 1, 40, generator=g) * 14000.0)
    _, theta_ref = O.oscillate_harmonics(f0, return_theta=True)
    theta, _ = N.harmonic_theta(f0[0, 0].numpy())
    assert np.array_equal(theta, theta_ref[0].numpy())


def test_noise_contract():
    g = torch.Generator().manual_seed(3)
    lf = 9
    kern = torch.rand(1, 961, lf, generator=g) + 0.5
    r01 = torch.rand(1, 961, lf, generator=g)
    ref = O.oscillate_noise(kern, r01)[0, 0].numpy()
    re, im = N.noise_spectrum(kern[0].numpy(), r01[0].numpy())
    got = N.noise_ola(re, im)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-6
    # the injected draw reproduces the reference's internal torch.rand under the same seed
    torch.manual_seed(11)
    r = torch.rand(1, 961, lf)
    torch.manual_seed(11)
    a = torch.rand(1, 961, lf) * 2 * np.pi - np.pi
    assert torch.equal(O.noise_angle(r), a)


def test_downsample_factors_pick_expected_taps():
    """A.1 consequences: /5 -> x[5d+2], /3 -> x[3d+1], /4 -> mean(x[4d+1], x[4d+2])."""
    x = np.arange(240, dtype=np.float32)[None]
    assert np.array_equal(N.interp_linear(x, 48, 1 / 5)[0], x[0, 2::5])
    assert np.array_equal(N.interp_linear(x, 80, 1 / 3)[0], x[0, 1::3])
    assert np.array_equal(N.interp_linear(x, 60, 1 / 4)[0], 0.5 * (x[0, 1::4] + x[0, 2::4]))


def test_sinc_resample_restates_torchaudio():
    """oracle/numerics_np.sinc_resample (the restated algorithm of torchaudio.functional.resample, reference infer.py:63-64)
    against outputs of torchaudio itself (tests/golden/resample.npz, made by make_golden.py golden_resample): the filter bank
    is bit-equal where torchaudio is importable; the resampled signals agree to fp32 accumulation error (torchaudio sums 171
    taps in fp32, the oracle in fp64)."""
    import math
    import numpy as np
    from conftest import load_golden
    from oracle import numerics_np as N
    g = load_golden("resample.npz")
    for sr, ch, n in g["cases"]:
        y = N.sinc_resample(g[f"in_{sr}"], int(sr), 24000)
        ref = g[f"out_{sr}"]
        assert y.shape == ref.shape == (ch, math.ceil(24000 * n / sr))
        assert float(np.abs(y - ref).max()) < 1e-5, sr
    try:
        import torchaudio.functional.functional as FF
    except Exception:
        return
    for sr in (44100, 48000, 16000, 22050, 8000, 32000, 11025):
        k, w = FF._get_sinc_resample_kernel(sr, 24000, math.gcd(sr, 24000))
        bank, width, o, nn = N.sinc_resample_bank(sr, 24000)
        assert w == width and np.array_equal(k[:, 0].numpy(), bank), sr

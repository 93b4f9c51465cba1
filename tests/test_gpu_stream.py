"""StreamInfer on the GPU: the SOLA kernel against the reference's SOLA logic on identical converted
windows, state handling over several ticks, and many concurrent streams."""
import pytest
import torch

from conftest import load_golden, t, rmse, max_abs
from oracle import tinyvc_oracle as O

pytestmark = pytest.mark.gpu


@torch.inference_mode()
def test_sola_matches_reference_logic(cuda_models, report):
    from tinyvc_b200.infer import Generator, StreamInfer
    enc, dec = cuda_models
    gen = Generator(enc, dec)
    g = load_golden("stream_4ticks.npz")
    index, blocks = t(g["index"]).cuda(), t(g["blocks"])
    si = StreamInfer(gen, target=index, pitch_shift=0.0, device=torch.device("cuda"))
    si.init_buffer()
    assert si.input_size == 13440
    captured = []

    def convert_fn(window):            # the oracle's SOLA runs on the windows the CUDA path converted
        return captured[-1]

    so = O.StreamOracle(convert_fn)
    torch.manual_seed(int(g["rand_seed"]))
    rands = [torch.rand(1, 961, 28) for _ in range(4)]
    worst = 0.0
    for i in range(4):
        real_convert = gen.convert

        def spy(wf, tgt, ps, *a, **k):
            y = real_convert(wf, tgt, ps, *a, **k)
            captured.append(y.cpu())
            return y

        gen.convert = spy
        try:
            out = si.audio_callback(blocks[i].cuda(), rand01=rands[i].cuda()).cpu()
        finally:
            gen.convert = real_convert
        ref = so.audio_callback(blocks[i].clone())
        assert int(si.last_shift[0]) == so.last_shift, f"tick {i}: SOLA shift {int(si.last_shift[0])} vs {so.last_shift}"
        worst = max(worst, max_abs(out, ref))
        assert out.shape == (1920,)
    report.add("stream_sola", max_abs=worst)
    assert worst < 1e-6
    report.add("stream_vs_golden_free_running", rmse=rmse(out, t(g["out_sola"])[3]))


@torch.inference_mode()
def test_batched_streams_match_single(cuda_models):
    from tinyvc_b200.infer import Generator, StreamInfer, BatchedStreamInfer
    enc, dec = cuda_models
    gen = Generator(enc, dec)
    gi = torch.Generator().manual_seed(3)
    index = torch.randn(1, 768, 64, generator=gi).cuda()
    S = 3
    blocks = 0.1 * torch.randn(2, S, 1920, generator=gi)
    rands = torch.rand(2, S, 961, 28, generator=gi)
    bs = BatchedStreamInfer(gen, S, target=index, device=torch.device("cuda"))
    bs.init_buffer()
    singles = [StreamInfer(gen, target=index, device=torch.device("cuda")) for _ in range(S)]
    for s in singles:
        s.init_buffer()
    for tick in range(2):
        out = bs.audio_callback(blocks[tick].cuda(), rand01=rands[tick].cuda())
        for s in range(S):
            one = singles[s].audio_callback(blocks[tick, s].cuda(), rand01=rands[tick, s:s + 1].cuda())
            assert torch.equal(one, out[s]), f"tick {tick} stream {s}"


@torch.inference_mode()
def test_output_pruning_does_not_change_a_tick(cuda_models):
    """audio_callback asks the decoder for the 5 760 samples SOLA reads (stream.py:75) instead of all 13 440: same blocks."""
    from tinyvc_b200.infer import Generator, BatchedStreamInfer
    enc, dec = cuda_models
    gen = Generator(enc, dec)
    gi = torch.Generator().manual_seed(11)
    index = torch.randn(1, 768, 2048, generator=gi).cuda()
    S = 4
    blocks = 0.1 * torch.randn(3, S, 1920, generator=gi)
    rands = torch.rand(3, S, 961, 28, generator=gi)
    a = BatchedStreamInfer(gen, S, target=index, device=torch.device("cuda"))
    b = BatchedStreamInfer(gen, S, target=index, device=torch.device("cuda"))
    b.prune_output = False
    a.init_buffer(); b.init_buffer()
    for tick in range(3):
        oa = a.audio_callback(blocks[tick].cuda(), rand01=rands[tick].cuda())
        ob = b.audio_callback(blocks[tick].cuda(), rand01=rands[tick].cuda())
        assert torch.equal(a.last_shift, b.last_shift)
        assert torch.equal(oa, ob), f"tick {tick}: max|d| = {float((oa - ob).abs().max()):.3e}"


@torch.inference_mode()
@pytest.mark.parametrize("pv", [False, True])
def test_tick_graph_replay_matches_eager_ticks(cuda_models, weights, pv):
    """The whole tick as one CUDA graph (window slide, analyse, retarget, synthesise, SOLA; state updated in place) against the
    eager launches, noise drawn in-kernel from identically seeded generators: same blocks bit for bit over several ticks, and
    a change of target re-captures."""
    from tinyvc_b200.infer import Generator, BatchedStreamInfer
    from tinyvc_b200.tinyvc import Decoder
    enc, dec_a = cuda_models
    dec_b = Decoder().eval()
    dec_b.load_state_dict(weights[1], strict=True)
    dec_b = dec_b.to("cuda")
    dec_a.seed_noise(99); dec_b.seed_noise(99)
    gi = torch.Generator().manual_seed(21)
    index = torch.randn(1, 768, 2048, generator=gi).cuda()
    index2 = torch.randn(1, 768, 1100, generator=gi).cuda()
    S = 3
    blocks = 0.1 * torch.randn(6, S, 1920, generator=gi)
    a = BatchedStreamInfer(Generator(enc, dec_a), S, target=index, device=torch.device("cuda"), use_phase_vocoder=pv)
    b = BatchedStreamInfer(Generator(enc, dec_b), S, target=index, device=torch.device("cuda"), use_phase_vocoder=pv)
    b.use_graph = False
    a.init_buffer(); b.init_buffer()
    for tick in range(6):
        if tick == 4:
            a.target = index2; b.target = index2
        oa = a.audio_callback(blocks[tick].cuda())
        ob = b.audio_callback(blocks[tick].cuda())
        assert torch.equal(a.last_shift, b.last_shift), f"tick {tick}"
        assert torch.equal(oa, ob), f"tick {tick}: max|d| = {float((oa - ob).abs().max()):.3e}"
        assert torch.equal(a.input_wav, b.input_wav) and torch.equal(a.sola_buffer, b.sola_buffer)
    assert a._graph is not None and b._graph is None
    # eager ticks (an injected draw) and replayed ticks alternate on the same state
    r = torch.rand(S, 961, 28, generator=gi).cuda()
    oa = a.audio_callback(blocks[0].cuda(), rand01=r)
    ob = b.audio_callback(blocks[0].cuda(), rand01=r)
    assert torch.equal(oa, ob)
    oa = a.audio_callback(blocks[1].cuda())
    ob = b.audio_callback(blocks[1].cuda())
    assert torch.equal(oa, ob) and torch.equal(a.input_wav, b.input_wav)


@torch.inference_mode()
def test_phase_vocoder_matches_reference(report):
    """phase_vocoder(a, b, fade_out, fade_in) (reference stream.py:9-26) for several stream pairs, n = 1920 and an
    odd length (the reference doubles a different bin range for odd n)."""
    from tinyvc_b200.infer.stream import phase_vocoder
    import math
    worst = 0.0
    for n, S, seed in ((1920, 3, 0), (1920, 1, 1), (481, 2, 2), (64, 1, 3)):
        g = torch.Generator().manual_seed(seed)
        fade_in = torch.sin(math.pi * torch.arange(0, 1, 1 / n)[:n] / 2) ** 2
        fade_out = 1 - fade_in
        a = 0.3 * torch.randn(S, n, generator=g)
        # b: a delayed, slightly different copy of a (what SOLA hands over) plus noise
        b = 0.9 * torch.roll(a, 7, dims=1) + 0.05 * torch.randn(S, n, generator=g)
        got = phase_vocoder(a.cuda(), b.cuda(), fade_out.cuda(), fade_in.cuda()).cpu()
        for s in range(S):
            want = O.phase_vocoder(a[s], b[s], fade_out, fade_in)
            worst = max(worst, max_abs(got[s], want))
    report.add("phase_vocoder", max_abs=worst)
    assert worst < 2e-5      # fp32 FFT vs direct DFT; a 2*pi phase-wrap flip would show as ~1e-3


@torch.inference_mode()
def test_stream_phase_vocoder_ticks(cuda_models, report):
    """use_phase_vocoder=True ticks against the oracle's SOLA + phase_vocoder on the windows the CUDA path converted."""
    from tinyvc_b200.infer import Generator, StreamInfer
    enc, dec = cuda_models
    gen = Generator(enc, dec)
    g = load_golden("stream_4ticks.npz")
    index, blocks = t(g["index"]).cuda(), t(g["blocks"])
    si = StreamInfer(gen, target=index, pitch_shift=0.0, device=torch.device("cuda"), use_phase_vocoder=True)
    si.init_buffer()
    captured = []
    so = O.StreamOracle(lambda w: captured[-1], use_phase_vocoder=True)
    torch.manual_seed(int(g["rand_seed"]))
    rands = [torch.rand(1, 961, 28) for _ in range(3)]
    real_convert = gen.convert

    def spy(wf, tgt, ps, *a, **k):
        y = real_convert(wf, tgt, ps, *a, **k)
        captured.append(y.cpu())
        return y

    gen.convert = spy
    worst = 0.0
    try:
        for i in range(3):
            out = si.audio_callback(blocks[i].cuda(), rand01=rands[i].cuda()).cpu()
            ref = so.audio_callback(blocks[i].clone())
            assert int(si.last_shift[0]) == so.last_shift
            if i == 0:
                # First tick: a = the all-zero initial sola_buffer, so fa = rfft(0) and angle(fa) is the angle of a
                # signed zero -- the reference gets 0 or +-pi per bin depending on which zeros its FFT library
                # returns as -0.0 (479 of 961 bins on torch 2.11 CPU), i.e. the cross-fade from silence is not a
                # function of the input.  Only the state hand-over is checked on this tick (tail = raw audio).
                assert torch.equal(si.sola_buffer[0].cpu(), so.sola_buffer)
                continue
            worst = max(worst, max_abs(out, ref))
    finally:
        gen.convert = real_convert
    report.add("stream_phase_vocoder", max_abs=worst)
    assert worst < 2e-5

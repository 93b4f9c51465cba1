"""Drop-in boundary checks that need no GPU: the C-ABI library loads and exports every symbol the
header declares, the Python mirror has the reference's state_dict keys / signatures, and there is no
silent CPU fallback."""
import inspect
import os
import re

import pytest
import torch

from conftest import REPO


def _header_symbols():
    txt = open(os.path.join(REPO, "include", "tinyvc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tvc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from tinyvc_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 25
    L = _lib.lib()
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/tinyvc_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == syms, "ctypes signature table out of sync with the header"
    assert b"sm_100a" in L.tvc_version()


def test_state_dict_contract(state_keys):
    from tinyvc_b200 import _lib
    from tinyvc_b200.tinyvc import Decoder, Encoder
    for kind, cls, kid in (("encoder", Encoder, 1), ("decoder", Decoder, 0)):
        sd = cls().state_dict()
        assert [[k, list(v.shape)] for k, v in sd.items()] == state_keys[kind], f"{kind} keys differ from the reference"
        assert tuple((k, v.numel()) for k, v in sd.items()) == _lib.param_names(kid)
        assert sum(v.numel() for v in sd.values()) == _lib.lib().tvc_param_total(kid)


def test_reference_import_paths_and_signatures():
    from module.tinyvc import Encoder, Decoder, match_features
    from module.infer import Generator, StreamInfer
    from module.utils import spectrogram, shift_frequency, estimate_energy, autopad_waveform
    assert list(inspect.signature(Decoder.__init__).parameters)[1:] == ["sample_rate", "n_fft", "frame_size", "num_harmonics"]
    assert list(inspect.signature(Encoder.__init__).parameters)[1:] == ["n_fft", "hop_size"]
    assert list(inspect.signature(match_features).parameters)[:5] == ["source", "reference", "k", "alpha", "metrics"]
    assert list(inspect.signature(Generator.convert).parameters)[1:6] == ["wf", "tgt", "pitch_shift", "f0_estimation", "device"]
    assert list(inspect.signature(StreamInfer.__init__).parameters)[1:] == [
        "generator", "target", "pitch_shift", "device", "block_size", "extra_size", "use_phase_vocoder", "f0_estimation"]
    d = Decoder()
    for attr in ("source_net", "filter_net", "frame_size", "sample_rate", "num_harmonics", "n_fft", "dsp", "infer"):
        assert hasattr(d, attr)
    e = Encoder()
    for attr in ("n_fft", "hop_size", "ssl_feature_estimator", "pitch_estimator"):
        assert hasattr(e, attr)
    for attr in ("freq2id", "id2freq", "decode", "num_classes"):
        assert hasattr(e.pitch_estimator, attr)
    assert autopad_waveform(torch.zeros(2, 700)).shape == (2, 960)
    assert autopad_waveform(torch.zeros(2, 960)).shape == (2, 960)


def test_no_cpu_fallback():
    from tinyvc_b200.tinyvc import Decoder, match_features
    from tinyvc_b200.utils import spectrogram
    from tinyvc_b200 import synth
    inp = synth.decoder_inputs(1, 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        Decoder().infer(inp["content"], inp["f0"], inp["energy"])
    with pytest.raises(RuntimeError, match="CUDA"):
        spectrogram(torch.zeros(1, 4800))
    with pytest.raises(RuntimeError, match="CUDA"):
        match_features(torch.zeros(1, 768, 4), torch.zeros(1, 768, 16))


def test_round2_host_logic_without_a_gpu():
    """Host-side pieces of the round-2 additions: the resampler's length contract (torchaudio: ceil(new * L / orig)), its refusal of
    CPU tensors, the sample range a streaming tick asks the decoder for (stream.py:75: y[-9600:-3840] of 13 440), and that
    the keep= range is validated before anything is launched."""
    import math
    from tinyvc_b200 import _lib
    from tinyvc_b200.infer import BatchedStreamInfer, StreamInfer
    from tinyvc_b200.utils import resample
    L = _lib.lib()
    for n, o, nw in [(4410, 44100, 24000), (4801, 48000, 24000), (1777, 16000, 24000), (7, 32000, 24000), (100, 24000, 24000)]:
        assert L.tvc_resample_length(n, o, nw) == math.ceil(nw * n / o)
    assert L.tvc_resample_length(-1, 44100, 24000) == -1 and L.tvc_resample_length(10, 0, 24000) == -1
    with pytest.raises(RuntimeError, match="CUDA"):
        resample(torch.zeros(1, 100), 44100, 24000)
    si = StreamInfer(None, device=torch.device("cuda"))
    assert si.input_size == 13440 and si._keep_range() == (3840, 9600)
    bs = BatchedStreamInfer(None, 4, block_size=960, extra_size=20000)     # window = block + extra, padded to whole frames
    assert bs.input_size == 20960 and bs._keep_range() == (21120 - 960 - 1920 - 1920 - 3840, 21120 - 3840)
    bs.prune_output = False
    assert bs._keep_range() is None
    with pytest.raises(RuntimeError):
        BatchedStreamInfer(None, 2, device=torch.device("cpu")).init_buffer()


def test_workspace_queries_are_monotone():
    from tinyvc_b200 import _lib
    L = _lib.lib()
    a, b, c = (L.tvc_decoder_workspace_bytes(*s) for s in ((1, 18), (4, 18), (4, 36)))
    assert 0 < a < b < c
    assert L.tvc_decoder_workspace_bytes(0, 5) == 0
    assert L.tvc_encoder_workspace_bytes(2, 10) > 0
    assert L.tvc_spectrogram_workspace_bytes(2, 4800) > 0 and L.tvc_spectrogram_workspace_bytes(2, 4801) == 0


def test_product_path_never_imports_oracle():
    """The oracle is test infrastructure: nothing under tinyvc_b200/ or module/ may reference it."""
    for root in ("tinyvc_b200", "module"):
        for dp, _, files in os.walk(os.path.join(REPO, root)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dp, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


def test_cli_flags_match_the_reference():
    """Flag spellings / defaults of reference infer.py:18-29 and infer_streaming.py:19-33 (device default is
    cuda here: there is no CPU path)."""
    import infer
    import infer_streaming
    a = vars(infer.build_parser().parse_args([]))
    assert a == dict(inputs="./inputs/", outputs="./outputs/", encoder_path="./models/encoder.pt",
                     decoder_path="./models/decoder.pt", f0_estimation="default", index="NONE", target="target.wav",
                     device="cuda", pitch_shift=0.0, chunk_size=1920, buffer_size=4, no_chunking=False)
    short = infer.build_parser().parse_args("-i a -o b -encp e -decp d -f0-est x -idx i -t t -d cuda:1 -p 2 -c 1 -b 2 -nc 1".split())
    assert (short.inputs, short.index, short.device, short.pitch_shift, short.no_chunking) == ("a", "i", "cuda:1", 2.0, True)
    b = vars(infer_streaming.build_parser().parse_args([]))
    b.pop("dry_run")
    assert b == dict(encoder_path="./models/encoder.pt", decoder_path="./models/decoder.pt", input=0, output=0,
                     loopback=-1, index="NONE", pitch_shift=0, target="target.wav", chunk=1920, extra=3840,
                     device="cuda", sample_rate=24000, input_gain=0, output_gain=0, f0_estimation="default")
    s = infer_streaming.build_parser().parse_args("-i 1 -o 2 -l 3 -c 960 -e 0 -sr 48000 -ig 3 -og -3 -f0-est dio".split())
    assert (s.input, s.output, s.loopback, s.chunk, s.extra, s.sample_rate, s.f0_estimation) == (1, 2, 3, 960, 0, 48000, "dio")
    with pytest.raises(SystemExit, match="no CPU path"):
        infer.load_generator("x", "y", torch.device("cpu"))

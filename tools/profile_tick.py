"""Developer: where a streaming tick goes (128 streams, window 13 440 samples): per-stage CUDA-event times and host time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import _lib
from tinyvc_b200.infer import BatchedStreamInfer, Generator
from tinyvc_b200.tinyvc import Decoder, Encoder, match_features
from tinyvc_b200.utils import estimate_energy, shift_frequency, spectrogram, autopad_waveform
from tinyvc_b200.weights import load_synth_weights

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
enc = load_synth_weights(Encoder().eval(), 7).to(dev)
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
gen = Generator(enc, dec)
g = torch.Generator(device=dev); g.manual_seed(3)
index = torch.randn(1, 768, 2048, device=dev, generator=g)
bs = BatchedStreamInfer(gen, S, target=index, device=dev)
bs.init_buffer()
blocks = 0.1 * torch.randn(S, 1920, device=dev, generator=g)

def ev(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return a.elapsed_time(b) / n, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3

print("tick  gpu %.3f ms  host-issue %.3f ms  wall %.3f ms" % ev(lambda: bs.audio_callback(blocks)))
wf = bs.input_wav
spec = spectrogram(wf); energy = estimate_energy(wf)
z, f0 = enc.infer(spec); zm = match_features(z, index); f0s = shift_frequency(f0, 0.0)
for name, fn in [("spectrogram", lambda: spectrogram(wf)), ("energy", lambda: estimate_energy(wf)), ("encoder", lambda: enc.infer(spec)),
                 ("knn", lambda: match_features(z, index)), ("shift", lambda: shift_frequency(f0, 0.0)),
                 ("decoder", lambda: dec.infer(zm, f0s, energy))]:
    print("%-12s gpu %.3f ms  host-issue %.3f ms  wall %.3f ms" % ((name,) + ev(fn)))
n0 = _lib.launch_count(); bs.audio_callback(blocks); print("launches per tick", _lib.launch_count() - n0)

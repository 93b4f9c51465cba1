"""Tiny end-to-end pass for compute-sanitizer (memcheck): every kernel once on small shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import synth
from tinyvc_b200.tinyvc import Decoder, Encoder, match_features
from tinyvc_b200.infer import Generator, StreamInfer
from tinyvc_b200.weights import load_synth_weights

dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
enc = load_synth_weights(Encoder().eval(), 7).to(dev)
inp = {k: v.to(dev) for k, v in synth.decoder_inputs(2, 3, 1).items()}
out = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
gen = Generator(enc, dec)
p = synth.pipeline_inputs(2, 2500, 37, 2)
y = gen.convert(p["wf"].to(dev), p["index"].to(dev), 1.0)
si = StreamInfer(gen, target=p["index"].to(dev), device=dev)
si.init_buffer()
o = si.audio_callback(torch.randn(1920, device=dev) * 0.1)
out2 = dec.infer(inp["content"], inp["f0"], inp["energy"])                     # in-kernel noise draw
big = synth.pipeline_inputs(1, 2400, 1100, 3)                                    # N >= 1024: tensor-core screened kNN
y2 = gen.convert(big["wf"].to(dev), big["index"].to(dev), 0.0)
sp = StreamInfer(gen, target=p["index"].to(dev), device=dev, use_phase_vocoder=True)
sp.init_buffer()
o2 = sp.audio_callback(torch.randn(1920, device=dev) * 0.1)
o2 = sp.audio_callback(torch.randn(1920, device=dev) * 0.1)
# round-2 additions: device resampler, pruned decoder range, split kNN sweep (2048-vector index, few queries), streaming ticks on
# the eager path with the shared-encoder launch shaping, several streams
from tinyvc_b200.utils import resample
from tinyvc_b200.infer import BatchedStreamInfer
r = resample(torch.randn(2, 1001, device=dev), 44100, 24000)
out3 = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"], keep=(500, 901))
idx2k = torch.randn(1, 768, 2048, device=dev)
zq = torch.randn(3, 768, 28, device=dev)
m2 = match_features(zq, idx2k)
bsi = BatchedStreamInfer(gen, 3, target=idx2k, device=dev)
bsi.use_graph = False
bsi.init_buffer()
for _ in range(2):
    o3 = bsi.audio_callback(torch.randn(3, 1920, device=dev) * 0.1)
torch.cuda.synchronize()
print("sanitize_small ok", out.shape, y.shape, o.shape, float(out.abs().max()), float(y.abs().max()))

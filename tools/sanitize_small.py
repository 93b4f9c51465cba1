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
torch.cuda.synchronize()
print("sanitize_small ok", out.shape, y.shape, o.shape, float(out.abs().max()), float(y.abs().max()))

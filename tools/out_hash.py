"""Developer: SHA-1 of Decoder.infer's waveform for fixed seeded inputs (compare two builds of the library: TVC_LIB=...)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import synth
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights
dec = load_synth_weights(Decoder().eval(), 7).to("cuda")
h = hashlib.sha1()
for B, Lf in [(3, 1), (5, 7), (64, 18), (2, 100)]:
    inp = {k: v.to("cuda") for k, v in synth.decoder_inputs(B, Lf, 4321 + Lf).items()}
    out = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
    h.update(out.cpu().numpy().tobytes())
print(os.environ.get("TVC_LIB", "in-tree"), h.hexdigest())

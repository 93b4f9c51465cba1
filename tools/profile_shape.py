"""Developer: per-launcher event profile of Decoder.infer at an arbitrary shape (default: one micro-batch of the config-4 share).
    python tools/profile_shape.py [B] [Lf]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import _lib, synth
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Lf = int(sys.argv[2]) if len(sys.argv) > 2 else 500
dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
g = torch.Generator(device=dev); g.manual_seed(5)
content = torch.randn(B, 768, Lf, device=dev, generator=g)
f0 = (220 * torch.exp2(torch.cumsum(0.03 * torch.randn(B, 1, Lf, device=dev, generator=g), 2))).clamp_(80, 800)
energy = torch.rand(B, 1, Lf * 480, device=dev, generator=g)
for _ in range(3):
    dec.infer(content, f0, energy)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    dec.infer(content, f0, energy)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
_lib.set_option("profile", "1")
for _ in range(2):
    dec.infer(content, f0, energy)
prof = _lib.profile_report()
_lib.set_option("profile", "0")
tot = sum(v["ms"] for v in prof.values()) / 2
print(f"B={B} Lf={Lf}: {ms:.3f} ms/step (graph)  {B*Lf*480/ms/1e6:.1f} M samples/s  frac {B*Lf*480/ms*1e3*3706.3/6551.7e9:.3f}; profile-pass sum {tot:.3f} ms")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:22s} {v['ms']/2*1e3:9.1f} us  x{v['launches']//2}  {v['ms']/2/tot*100:5.1f} %")

"""Developer: three streaming ticks (128 streams) and nothing else -- the command an ncu launch list of one tick is taken from."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200.infer import BatchedStreamInfer, Generator
from tinyvc_b200.tinyvc import Decoder, Encoder
from tinyvc_b200.weights import load_synth_weights

dev = torch.device("cuda:0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
gen = Generator(load_synth_weights(Encoder().eval(), 7).to(dev), load_synth_weights(Decoder().eval(), 7).to(dev))
g = torch.Generator(device=dev); g.manual_seed(3)
index = torch.randn(1, 768, 2048, device=dev, generator=g)
bs = BatchedStreamInfer(gen, S, target=index, device=dev)
bs.init_buffer()
blocks = 0.1 * torch.randn(S, 1920, device=dev, generator=g)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    bs.audio_callback(blocks)
torch.cuda.synchronize()

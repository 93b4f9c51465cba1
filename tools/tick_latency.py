"""Developer: tick time of S streams with the tick replayed as one CUDA graph vs launched eagerly (wall clock per tick with a
synchronise after every tick = what a real-time caller waits for, and back-to-back throughput)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import _lib
from tinyvc_b200.infer import BatchedStreamInfer, Generator
from tinyvc_b200.tinyvc import Decoder, Encoder
from tinyvc_b200.weights import load_synth_weights

dev = torch.device("cuda:0")
gen = Generator(load_synth_weights(Encoder().eval(), 7).to(dev), load_synth_weights(Decoder().eval(), 7).to(dev))
g = torch.Generator(device=dev); g.manual_seed(3)
index = torch.randn(1, 768, 2048, device=dev, generator=g)
for S in [int(a) for a in sys.argv[1:]] or [1, 8, 128]:
    blocks = 0.1 * torch.randn(S, 1920, device=dev, generator=g)
    for graph in (False, True):
        bs = BatchedStreamInfer(gen, S, target=index, device=dev)
        bs.use_graph = graph
        bs.init_buffer()
        for _ in range(4):
            bs.audio_callback(blocks)
        torch.cuda.synchronize()
        lat = []
        for _ in range(30):
            t0 = time.perf_counter(); bs.audio_callback(blocks); torch.cuda.synchronize(); lat.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _ in range(30):
            bs.audio_callback(blocks)
        torch.cuda.synchronize()
        thr = (time.perf_counter() - t0) / 30
        n0 = _lib.launch_count(); bs.audio_callback(blocks); torch.cuda.synchronize()
        lat.sort()
        print(f"S={S:4d} graph={int(graph)}: latency median {lat[15]*1e3:.3f} ms  p90 {lat[27]*1e3:.3f} ms   back-to-back {thr*1e3:.3f} ms/tick")

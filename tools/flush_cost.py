"""Developer: config-2 decoder step with and without the 512 MB L2 flush between steps (what cold weights / inputs cost).
Measured on B200: 0.932 ms with the flush, 0.929 ms without -- cold L2 is not what bounds the step, so prefetching the weights
into L2 at the start of a step (round 1 tried it as a serial kernel) has nothing to win."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import synth
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights
dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
inp = {k: v.to(dev) for k, v in synth.decoder_inputs(64, 18, 1236).items()}
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
step = lambda: dec.infer(inp["content"], inp["f0"], inp["energy"])
for _ in range(5): step()
torch.cuda.synchronize()
for do_flush in (True, False, True, False):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(40)]
    for a, b in evs:
        if do_flush: flush.fill_(1)
        a.record(); step(); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)
    print(f"flush={int(do_flush)}: mean {sum(ms)/len(ms):.4f} ms  median {ms[len(ms)//2]:.4f} ms")

"""Developer timing probe (not the contract bench): decoder stages at BASELINE config 2 on cuda:0."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import synth, _lib
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights

B, LF = int(os.environ.get("QB_B", 64)), int(os.environ.get("QB_LF", 18))
dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
inp = {k: v.to(dev) for k, v in synth.decoder_inputs(B, LF, 1236).items()}


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


amps, kern = dec.source_net(inp["content"], inp["f0"], inp["energy"])
src = dec.dsp(inp["f0"], amps, kern, rand01=inp["rand01"])
t_all = timeit(lambda: dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"]))
t_sn = timeit(lambda: dec.source_net(inp["content"], inp["f0"], inp["energy"]))
t_dsp = timeit(lambda: dec.dsp(inp["f0"], amps, kern, rand01=inp["rand01"]))
t_fn = timeit(lambda: dec.filter_net(inp["content"], inp["f0"], inp["energy"], src))
n = B * LF * 480
print(f"B={B} Lf={LF} samples={n}")
print(f"decoder.infer {t_all:.3f} ms -> {n / t_all / 1e3:.1f} M samples/s; conv1d-hbm-frac {n / t_all * 1e3 * 3706.3 / 6534.8e9:.4f}")
print(f"source_net {t_sn:.3f} ms | dsp {t_dsp:.3f} ms | filter_net {t_fn:.3f} ms")

"""Developer: record the per-role pipeline timeline of selected tc_conv launches of one Decoder.infer step (config 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import _lib, synth
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights

which = sys.argv[1] if len(sys.argv) > 1 else "23,43,44,47"
out = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/tc_trace.txt"
dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
inp = {k: v.to(dev) for k, v in synth.decoder_inputs(64, 18, 1236).items()}
_lib.set_option("graphs", "0")
for _ in range(3):
    dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
torch.cuda.synchronize()
_lib.set_option("tc_trace", which)
dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
torch.cuda.synchronize()
_lib.set_option("tc_trace_dump", out)
print("wrote", out, os.path.getsize(out))

"""Minimal driver for profiler runs: N decoder steps at BASELINE config 2 (B=64, Lf=18), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import synth
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights
B, LF, N = int(os.environ.get("QB_B", 64)), int(os.environ.get("QB_LF", 18)), int(os.environ.get("QB_N", 3))
dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
inp = {k: v.to(dev) for k, v in synth.decoder_inputs(B, LF, 1236).items()}
for _ in range(N):
    dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
torch.cuda.synchronize()

"""Developer: config-4 share (512 x 500 frames in micro-batches of 64): loop over 8 different slices vs 8 x the same slice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import device_decoder_inputs, time_events, C4_SHARE, C4_LF, C4_MB, FRAME
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights
dev = torch.device("cuda:0")
dec = load_synth_weights(Decoder().eval(), 7).to(dev)
inp = device_decoder_inputs(C4_SHARE, C4_LF, dev, 1238)
res = torch.empty(C4_SHARE, C4_LF * FRAME, device=dev)
def c4():
    for o in range(0, C4_SHARE, C4_MB):
        dec.infer(inp["content"][o:o + C4_MB], inp["f0"][o:o + C4_MB], inp["energy"][o:o + C4_MB], out=res[o:o + C4_MB])
def same():
    for o in range(0, C4_SHARE, C4_MB):
        dec.infer(inp["content"][:C4_MB], inp["f0"][:C4_MB], inp["energy"][:C4_MB], out=res[:C4_MB])
for mb in (64, 32, 128):
    C4_MB = mb
    print("mb", mb, "8 slices: %.2f ms" % time_events(c4, 3, 3), " same slice: %.2f ms" % time_events(same, 3, 3), flush=True)
import subprocess, threading, time
samples = []
stop = False
def poll():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.1)
C4_MB = 64
th = threading.Thread(target=poll); th.start()
t0 = time.time()
ms = time_events(c4, 15, 2)
stop = True; th.join()
print("15 x c4: %.2f ms each" % ms)
print("clock samples:", samples[:3], "...", samples[len(samples)//2], "...", samples[-2:])

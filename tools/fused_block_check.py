"""Developer check of the experimental fused Upsample block (tvc_set_option("fused_up", "1"), csrc/tc_block.cu).

Runs decoder.infer on cuda:0 with the five separate conv launches of ups.4 and with the fused kernel on the same seeded
inputs and noise draw, and reports whether the waveforms are bit-identical (they are meant to be: same MMA order per
tile, same epilogue arithmetic) plus the step time of each.  Not part of tests/ until the kernel has been validated:
a protocol error in it traps the launch, which would take the rest of a pytest process with it.

    python tools/fused_block_check.py            # configs: (B, Lf) = (2, 1), (3, 2), (64, 18)
    FB_CASES="4x5,64x18" python tools/fused_block_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tinyvc_b200 import _lib, synth
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights


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


def main() -> int:
    dev = torch.device("cuda:0")
    dec = load_synth_weights(Decoder().eval(), 7).to(dev)
    cases = [tuple(int(v) for v in c.split("x")) for c in os.environ.get("FB_CASES", "2x1,3x2,64x18").split(",")]
    bad = 0
    for B, Lf in cases:
        inp = {k: v.to(dev) for k, v in synth.decoder_inputs(B, Lf, 1236).items()}
        run = lambda: dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
        _lib.set_option("fused_up", "0")
        ref = run().clone()
        t_ref = timeit(run)
        _lib.set_option("fused_up", "1")
        got = run().clone()
        t_got = timeit(run)
        _lib.set_option("fused_up", "0")
        diff = (got - ref).abs()
        same = bool(torch.equal(got, ref))
        nz = int((diff > 0).sum())
        first = int(torch.nonzero(diff.flatten() > 0)[0]) if nz else -1
        print(f"B={B} Lf={Lf}: bit-identical={same} max|d|={float(diff.max()):.3e} differing={nz}/{diff.numel()} first={first} "
              f"(t={first % (Lf * 480) if nz else -1}) | separate {t_ref:.3f} ms, fused {t_got:.3f} ms", flush=True)
        bad += 0 if float(diff.max()) < 2e-5 else 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

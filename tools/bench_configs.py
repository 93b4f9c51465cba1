#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs at their per-GPU sizes (not the contract line; bench.py is).

    python tools/bench_configs.py [--configs 3,4,5] [--steps K]

  3: full Encoder -> kNN(50k index) -> Decoder, batch 256 x 4 s        (samples/s of converted audio)
  4: Decoder batch 512 x 10 s (the per-GPU share of 4096 over 8 GPUs)  (samples/s)
  5: 128 concurrent streams, one tick = 1920 new samples per stream    (samples/s, ticks/s)
One JSON line per config; CUDA events, inputs resident in HBM, 3 warm-ups.
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from tinyvc_b200 import _lib, synth  # noqa: E402
from tinyvc_b200.infer import BatchedStreamInfer, Generator  # noqa: E402
from tinyvc_b200.tinyvc import Decoder, Encoder  # noqa: E402
from tinyvc_b200.weights import load_synth_weights  # noqa: E402


def timed(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, (_lib.launch_count() - n0) // steps


def stage_times(gen, wf, index, rand01, steps=3):
    """Per-stage CUDA-event times of Generator.convert (front end, encoder, kNN, decoder)."""
    from tinyvc_b200.tinyvc import match_features
    from tinyvc_b200.utils import estimate_energy, shift_frequency, spectrogram
    res = {}

    def t(name, fn):
        ms, _ = timed(fn, steps, warm=1)
        res[name] = round(ms, 3)

    spec = spectrogram(wf)
    energy = estimate_energy(wf)
    z, f0 = gen.encoder.infer(spec)
    zm = match_features(z, index)
    t("spectrogram+energy", lambda: (spectrogram(wf), estimate_energy(wf)))
    t("encoder", lambda: gen.encoder.infer(spec))
    t("knn", lambda: match_features(z, index))
    t("decoder", lambda: gen.decoder.infer(zm, shift_frequency(f0, 0.0), energy, rand01=rand01))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    enc = load_synth_weights(Encoder().eval(), seed=7).to(dev)
    dec = load_synth_weights(Decoder().eval(), seed=7).to(dev)
    gen = Generator(enc, dec)
    for c in args.configs.split(","):
        c = int(c)
        if c == 3:
            B, T, N = 256, 96000, 50000
            inp = synth.pipeline_inputs(B, T, N, seed=1234 + 3)
            wf, index = inp["wf"].to(dev), inp["index"].to(dev)
            rand01 = torch.rand(B, 961, T // 480, device=dev)
            ms, launches = timed(lambda: gen.convert(wf, index, 0.0, rand01=rand01), args.steps)
            line = {"config": 3, "workload": "Encoder->kNN(50k)->Decoder, batch 256 x 4 s", "ms_per_step": ms,
                    "samples_per_s": B * T / ms * 1e3, "gpu_launches": launches, "stages_ms": stage_times(gen, wf, index, rand01)}
        elif c == 4:
            B, Lf = 512, 500
            inp = {k: v.to(dev) for k, v in synth.decoder_inputs(B, Lf, seed=1234 + 4).items()}
            ms, launches = timed(lambda: dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"]), args.steps)
            sps = B * Lf * 480 / ms * 1e3
            line = {"config": 4, "workload": "Decoder batch 512 x 10 s (per-GPU share of 4096 over 8 GPUs)", "ms_per_step": ms,
                    "samples_per_s": sps, "gpu_launches": launches, "hbm_conv1d_fraction": sps * 3706.3 / 6551.7e9}
        elif c == 5:
            S = 128
            index = torch.randn(1, 768, 2048, device=dev)
            bs = BatchedStreamInfer(gen, S, target=index, device=dev)
            bs.init_buffer()
            blocks = 0.1 * torch.randn(S, 1920, device=dev)
            rand01 = torch.rand(S, 961, 28, device=dev)
            ms, launches = timed(lambda: bs.audio_callback(blocks, rand01=rand01), args.steps * 4)
            line = {"config": 5, "workload": "128 streams, 13 440-sample window, 1 920 new samples per stream per tick",
                    "ms_per_tick": ms, "ticks_per_s": 1e3 / ms, "samples_per_s": S * 1920 / ms * 1e3, "gpu_launches": launches,
                    "realtime_streams_supported": S * 80.0 / ms}
        else:
            continue
        print(json.dumps(line), flush=True)
        torch.cuda.empty_cache()
        _lib.WORKSPACE.clear()


if __name__ == "__main__":
    main()

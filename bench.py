#!/usr/bin/env python
"""Contract benchmark: decoder audio samples/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N = 1): BASELINE.json configs[1] -- Decoder-only forward, batch = 64, 8192-sample chunks
(autopadded to 18 frames = 8640 samples), seeded random F0 + content, random
weights of the reference's architecture.  N > 1: every rank runs the same per-GPU batch on its own
utterances (utterances are independent; no data-path collective) -> "scaling": "weak"; in the same run the
north-star's multi-GPU mode is timed as well (`scatter_gather`): rank 0 holds BASELINE configs[3]'s job
(512 x N utterances of 10 s), `tinyvc_b200.shard.ShardedDecoder` shards it over the GPUs and returns the
waveforms to rank 0, with its efficiency against N x the single-GPU rate of the same share.

A step = one `Decoder.infer` over the batch: SourceNet -> harmonic+noise dsp -> FilterNet.
  value : samples/s with inputs resident in HBM, CUDA events, L2 flushed before every timed step,
          max over ranks.
  e2e   : the same step through the public Python API from pinned HOST buffers: H2D of content /
          f0 / energy and D2H of the waveform inside the timed region.
  roofline / cpu_baseline : see DESIGN.md "Measurement".
`--impl reference` times the reference's CPU algorithm (oracle port of module/tinyvc/decoder.py on
torch-CPU with all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

FRAME = 480
BATCH, LF = 64, 18                       # configs[1]: 64 x 8192-sample chunks -> 18 frames each
CONV1D_BYTES_PER_SAMPLE = 3706.3         # every Conv1d reads its input + writes its output once
COMPULSORY_BYTES_PER_SAMPLE = 22.4       # content 6.4 + energy 4 + out 4 + noise draw 8.0 (+f0)
FLOP_PER_SAMPLE = 105232.0               # 2 x 52 616 MAC
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12   # CUDA-core FMA peak at max clock; the measured one is used
C4_SHARE, C4_LF, C4_MB = 512, 500, 64    # configs[3]: 4096 x 10 s over 8 GPUs = 512 utterances of 500 frames per GPU
METRIC = "decoder audio samples/sec (RTF) at 1/2/4/8 B200 vs host-CPU reference"
WORKLOAD = "Decoder-only fwd, batch=64, 8192-sample chunks (18 frames = 8640 samples), random F0+content"


def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def measured_peak_hbm():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_decoder_rate(sd, inp, utts: int, reps: int, threads: int):
    """samples/s of the CPU oracle decoder on the first `utts` utterances of the workload."""
    from oracle import tinyvc_oracle as O
    torch.set_num_threads(threads)
    sel = slice(0, utts)
    args = (inp["content"][sel], inp["f0"][sel], inp["energy"][sel], inp["rand01"][sel])
    with torch.inference_mode():
        O.decoder_infer(sd, *args)                       # warm-up (first istft builds FFT plans)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            O.decoder_infer(sd, *args)
            ts.append(time.perf_counter() - t0)
    n = utts * LF * FRAME
    return n / statistics.median(ts), ts


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tinyvc_b200 import synth
    from tinyvc_b200.weights import synth_state_dict
    keys = json.load(open(os.path.join(REPO, "tests", "golden", "state_keys.json")))["decoder"]
    sd = synth_state_dict({k: torch.empty(shape) for k, shape in keys}, 7)   # reference's own key/shape list
    inp = synth.decoder_inputs(BATCH, LF, seed=1234 + 2)
    threads = os.cpu_count() or 1
    utts = 16
    from oracle import tinyvc_oracle as O
    torch.set_num_threads(threads)
    a = (inp["content"][:utts], inp["f0"][:utts], inp["energy"][:utts], inp["rand01"][:utts])
    with torch.inference_mode():
        for _ in range(max(1, min(args.warmup, 3))):
            O.decoder_infer(sd, *a)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.decoder_infer(sd, *a)
        dt = time.perf_counter() - t0
    n = utts * LF * FRAME
    val = n * args.steps / dt
    sample = f"{utts} of the {BATCH} utterances per step ({n} samples), oracle port of decoder.py on torch-CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random F0/content/weights)",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample,
                         "cpu": cpu_model()},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rtf_x": val / 24000.0,
    }
    print(json.dumps(line), flush=True)


def device_decoder_inputs(n: int, lf: int, dev, seed: int):
    """Synthetic decoder inputs of SURVEY 8(d)'s distributions generated ON the device (the 4096 x 10 s job is 10 GB:
    too slow to draw on the host): content ~ N(0,1), f0 random-walk contour 80..800 Hz with ~25 % unvoiced frames in
    runs, energy ~ U(0,1)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    content = torch.randn(n, 768, lf, device=dev, generator=g)
    walk = torch.cumsum(0.03 * torch.randn(n, lf, device=dev, generator=g), dim=1)
    f0 = (220.0 * torch.exp2(walk)).clamp_(80.0, 800.0)
    period = torch.randint(60, 120, (n, 1), device=dev, generator=g)
    phase = torch.randint(0, 120, (n, 1), device=dev, generator=g)
    t = torch.arange(lf, device=dev)[None, :]
    f0 = torch.where(((t + phase) % period) < period // 4, torch.zeros_like(f0), f0).unsqueeze(1).contiguous()
    energy = torch.rand(n, 1, lf * FRAME, device=dev, generator=g)
    return {"content": content, "f0": f0, "energy": energy}


def time_events(fn, k: int, warm: int) -> float:
    """ms per call: CUDA events on the current stream around k calls after `warm` untimed ones."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


def other_configs(dec, dev, peak_hbm: float) -> dict:
    """BASELINE.json configs[2], the per-GPU share of configs[3] and of configs[4] on this GPU (inputs resident in HBM,
    CUDA events, 2 warm-ups; working sets are far larger than L2)."""
    from tinyvc_b200 import _lib, synth
    from tinyvc_b200.infer import BatchedStreamInfer, Generator
    from tinyvc_b200.tinyvc import Encoder, match_features
    from tinyvc_b200.utils import estimate_energy, shift_frequency, spectrogram
    from tinyvc_b200.weights import load_synth_weights
    out = {}
    enc = load_synth_weights(Encoder().eval(), seed=7).to(dev)
    gen = Generator(enc, dec)
    # ---- configs[2]: Encoder -> kNN(50k) -> Decoder, 256 x 4 s
    B, T, N = 256, 96000, 50000
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + 3)
    wf = 0.1 * torch.randn(B, T, device=dev, generator=g)
    index = torch.randn(1, 768, N, device=dev, generator=g)
    ms = time_events(lambda: gen.convert(wf, index, 0.0), 3, 2)
    spec, energy = spectrogram(wf), estimate_energy(wf)
    z, f0 = enc.infer(spec)
    zm = match_features(z, index)
    stages = {"spectrogram+energy": time_events(lambda: (spectrogram(wf), estimate_energy(wf)), 3, 1),
              "encoder": time_events(lambda: enc.infer(spec), 3, 1),
              "knn": time_events(lambda: match_features(z, index), 3, 1),
              "decoder": time_events(lambda: dec.infer(zm, shift_frequency(f0, 0.0), energy), 3, 1)}
    sps = B * T / ms * 1e3
    dec_sps = B * T / stages["decoder"] * 1e3
    out["config3_full_pipeline"] = {
        "workload": "Encoder -> kNN(50k-vector index) -> Decoder, batch 256 x 4 s clips, 1 GPU", "ms_per_step": ms,
        "samples_per_s": sps, "stages_ms": {k: round(v, 3) for k, v in stages.items()},
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak_hbm, "achieved": dec_sps * CONV1D_BYTES_PER_SAMPLE / 1e9,
                     "frac": dec_sps * CONV1D_BYTES_PER_SAMPLE / (peak_hbm * 1e9), "of": "decoder stage, Conv1d-layer model"}}
    del wf, index, spec, energy, z, f0, zm
    torch.cuda.empty_cache()
    # ---- configs[3] share: Decoder 512 x 10 s
    inp = device_decoder_inputs(C4_SHARE, C4_LF, dev, 1234 + 4)
    res = torch.empty(C4_SHARE, C4_LF * FRAME, device=dev)

    def c4():
        for o in range(0, C4_SHARE, C4_MB):
            dec.infer(inp["content"][o:o + C4_MB], inp["f0"][o:o + C4_MB], inp["energy"][o:o + C4_MB], out=res[o:o + C4_MB])

    c4()
    cs = ClockSampler(dev.index or 0)
    ms = time_events(c4, 6, 1)
    clk4 = cs.stop()
    sps = C4_SHARE * C4_LF * FRAME / ms * 1e3
    out["config4_share"] = {
        "workload": "Decoder batch 512 x 10 s (one GPU's share of 4096 over 8 GPUs), micro-batches of 64", "ms_per_step": ms,
        "samples_per_s": sps, "clocks": clk4,
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak_hbm, "achieved": sps * CONV1D_BYTES_PER_SAMPLE / 1e9,
                     "frac": sps * CONV1D_BYTES_PER_SAMPLE / (peak_hbm * 1e9), "of": "whole step, Conv1d-layer model"}}
    del inp, res
    torch.cuda.empty_cache()
    # ---- configs[4] share: 128 concurrent streams, one tick = 1920 new samples per stream
    S = 128
    index = torch.randn(1, 768, 2048, device=dev, generator=g)
    bs = BatchedStreamInfer(gen, S, target=index, device=dev)
    bs.init_buffer()
    blocks = 0.1 * torch.randn(S, 1920, device=dev, generator=g)
    ms = time_events(lambda: bs.audio_callback(blocks), 20, 3)        # from the second tick on: one CUDA-graph replay per tick
    # kernels inside a tick: counted on an eager twin (a graph replay launches no kernel from the host)
    eager = BatchedStreamInfer(gen, S, target=index, device=dev)
    eager.use_graph = False
    eager.init_buffer()
    for _ in range(2):
        eager.audio_callback(blocks)
    n0 = _lib.launch_count()
    eager.audio_callback(blocks)
    kernels = _lib.launch_count() - n0
    ms_eager = time_events(lambda: eager.audio_callback(blocks), 10, 1)
    # what a single real-time stream waits for: wall clock of one tick including the synchronisation (the reference's use case)
    one = BatchedStreamInfer(gen, 1, target=index, device=dev)
    one.init_buffer()
    for _ in range(4):
        one.audio_callback(blocks[:1])
    torch.cuda.synchronize()
    lat = []
    for _ in range(30):
        t0 = time.perf_counter()
        one.audio_callback(blocks[:1])
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    out["config5_share"] = {
        "workload": "128 concurrent streams (one GPU's share of 1024), 13 440-sample window, 1 920 new samples per stream per tick",
        "ms_per_tick": ms, "samples_per_s": S * 1920 / ms * 1e3, "realtime_streams_supported": S * 80.0 / ms,
        "tick_mode": "whole tick replayed as one CUDA graph; the decoder computes only the 5 760 samples the SOLA step reads",
        "gpu_kernels_per_tick": kernels, "ms_per_tick_eager_launches": ms_eager,
        "single_stream_tick_latency_ms": {"median": lat[len(lat) // 2], "p90": lat[int(len(lat) * 0.9)],
                                          "audio_ms_per_tick": 80.0}}
    _lib.WORKSPACE.clear()
    torch.cuda.empty_cache()
    return out


def scatter_gather(dec, dev, rank: int, world: int, steps: int, barrier) -> dict:
    """North-star multi-GPU mode at BASELINE configs[3] size: rank 0 holds 512 x N utterances of 10 s; ShardedDecoder
    shards them over the N GPUs and the waveforms come back to rank 0.  Timed with CUDA events on every rank's stream
    between barriers, max over ranks; the single-GPU rate of one share (same micro-batching) is measured in the same run
    on rank 0 so that efficiency = aggregate / (N x single)."""
    import torch.distributed as dist
    from tinyvc_b200 import synth
    from tinyvc_b200.shard import ShardedDecoder
    sd = ShardedDecoder(dec, dev, micro_batch=C4_MB)
    n = C4_SHARE * world
    full = device_decoder_inputs(n, C4_LF, dev, 4321) if rank == 0 else None

    def step():
        return sd.infer(full["content"], full["f0"], full["energy"]) if rank == 0 else sd.infer()

    for _ in range(2):
        step()
    barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / steps
    samples = n * C4_LF * FRAME
    res = {"workload": f"Decoder batch {n} x 10 s held by rank 0, sharded over {world} GPUs ({C4_SHARE} per GPU), micro-batches of {C4_MB}",
           "transport": sd.transport, "value": samples / ms * 1e3, "unit": "samples/s", "ms_per_step": ms, "steps": steps,
           "utterances": n,
           "input_bytes_leaving_rank0_per_step": (world - 1) * C4_SHARE * (768 * C4_LF + C4_LF + C4_LF * FRAME) * 4,
           "output_bytes_entering_rank0_per_step": (world - 1) * C4_SHARE * C4_LF * FRAME * 4,
           "timing": "CUDA events on each rank's stream around K sharded calls between barriers, max over ranks; "
                     "transfers and both barriers of every call are inside"}
    # single-GPU rate of one share, same run, same micro-batching (rank 0, everyone else idle)
    if rank == 0:
        res1 = torch.empty(C4_SHARE, C4_LF * FRAME, device=dev)

        def one():
            for o in range(0, C4_SHARE, C4_MB):
                dec.infer(full["content"][o:o + C4_MB], full["f0"][o:o + C4_MB], full["energy"][o:o + C4_MB], out=res1[o:o + C4_MB])

        ms1 = time_events(one, max(2, steps // 2), 1)
        single = C4_SHARE * C4_LF * FRAME / ms1 * 1e3
        res.update(single_gpu_share={"ms_per_step": ms1, "samples_per_s": single},
                   speedup_vs_one_gpu=res["value"] / single, efficiency=res["value"] / (world * single))
        del res1
    barrier()
    # parity of the sharded path: a small job with an injected noise draw must equal rank 0 converting it alone, bit for bit
    small = None
    if rank == 0:
        small = {k: v.to(dev) for k, v in synth.decoder_inputs(3 * world + 1, 20, seed=99).items()}
    sdp = ShardedDecoder(dec, dev, micro_batch=2, transport=sd.transport)
    got = sdp.infer(small["content"], small["f0"], small["energy"], small["rand01"]) if rank == 0 else sdp.infer()
    if rank == 0:
        want = dec.infer(small["content"], small["f0"], small["energy"], rand01=small["rand01"])
        res["bit_identical_to_one_gpu"] = bool(torch.equal(got, want))
    barrier()
    sdp.close()
    sd.close()
    del full
    torch.cuda.empty_cache()
    return res


def streaming_share(dec, dev, rank: int, world: int, barrier) -> dict:
    """BASELINE configs[4] at N GPUs: 128 x N concurrent streams, every stream pinned to one rank for life
    (`shard.stream_owner`: its window and SOLA tail live there), one tick = 1 920 new samples per stream through the whole
    Encoder -> kNN -> Decoder -> SOLA chain.  No data-path exchange between ranks; ticks are timed with CUDA events between
    barriers, max over ranks."""
    import torch.distributed as dist
    from tinyvc_b200.infer import BatchedStreamInfer, Generator
    from tinyvc_b200.shard import stream_owner
    from tinyvc_b200.tinyvc import Encoder
    from tinyvc_b200.weights import load_synth_weights
    S = 128
    total = S * world
    mine = [i for i in (rank * S, rank * S + S - 1) if stream_owner(i, total, world) == rank]
    assert len(mine) == 2, "stream ownership does not match the contiguous partition"
    enc = load_synth_weights(Encoder().eval(), seed=7).to(dev)
    gen = Generator(enc, dec)
    g = torch.Generator(device=dev)
    g.manual_seed(99)                                   # the index is replicated: same seed on every rank
    index = torch.randn(1, 768, 2048, device=dev, generator=g)
    g.manual_seed(100 + rank)
    blocks = 0.1 * torch.randn(S, 1920, device=dev, generator=g)
    bs = BatchedStreamInfer(gen, S, target=index, device=dev)
    bs.init_buffer()
    for _ in range(3):
        bs.audio_callback(blocks)
    barrier()
    torch.cuda.synchronize()
    k = 20
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        bs.audio_callback(blocks)
    b.record()
    torch.cuda.synchronize()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / k
    return {"workload": f"{total} concurrent streams over {world} GPUs ({S} per GPU, pinned), 1 920 new samples per stream per tick",
            "ms_per_tick": ms, "samples_per_s": total * 1920 / ms * 1e3, "realtime_streams_supported": total * 80.0 / ms,
            "timing": "CUDA events around 20 ticks on every rank between barriers, max over ranks"}


def run_ours(args) -> None:
    import torch.distributed as dist
    from tinyvc_b200 import _lib, synth
    from tinyvc_b200.tinyvc import Decoder
    from tinyvc_b200.weights import load_synth_weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        # Keep stdout to the one JSON line: NCCL prints its version banner (and NCCL_DEBUG output) on stdout when the
        # first communicator is created, so file descriptor 1 points at stderr until that has happened.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()

    dec = load_synth_weights(Decoder().eval(), seed=7)
    sd_cpu = {k: v.clone() for k, v in dec.state_dict().items()} if rank == 0 else None
    dec = dec.to(dev)
    host = synth.decoder_inputs(BATCH, LF, seed=1234 + 2 + rank)     # every rank its own utterances
    inp = {k: v.to(dev) for k, v in host.items()}
    out_pinned = torch.empty(BATCH, LF * FRAME).pin_memory()
    samples = BATCH * LF * FRAME
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    # The public call, as the reference's callers make it (infer.py:66 -> decoder.infer(content, f0, energy)): the
    # uniform noise draw of decoder.py:78 happens inside the step (in the noise kernel; torch.rand in the reference).
    dec.seed_noise(1234 + rank)

    def step():
        return dec.infer(inp["content"], inp["f0"], inp["energy"])

    # e2e staging: the three inputs are views of ONE pinned host buffer and ONE device buffer, so a step's host -> device
    # traffic is a single copy (three separate copies cost ~10 us of launch / completion overhead each on a ~1 ms step)
    names = ("content", "f0", "energy")
    sizes = [host[k].numel() for k in names]
    offs = [sum(-(-n // 64) * 64 for n in sizes[:i]) for i in range(3)]            # 256-byte aligned views
    total = offs[2] + sizes[2]
    stage_host = torch.empty(total, dtype=torch.float32).pin_memory()
    stage_dev = torch.empty(total, dtype=torch.float32, device=dev)
    e2e_dev = {}
    for k, o, n in zip(names, offs, sizes):
        stage_host[o:o + n].copy_(host[k].reshape(-1))
        e2e_dev[k] = stage_dev[o:o + n].view(host[k].shape)

    def step_e2e():
        # host -> device: the step's inputs from pinned memory; device -> host: the waveform goes straight into the pinned
        # result buffer (Decoder.infer(out=pinned): the last kernel stores over PCIe, no separate copy)
        stage_dev.copy_(stage_host, non_blocking=True)
        return dec.infer(e2e_dev["content"], e2e_dev["f0"], e2e_dev["energy"], out=out_pinned)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None

    def timed(fn, k):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        barrier()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        for a, b in evs:
            flush.fill_(1)                      # evict L2 between timed iterations (not timed)
            a.record()
            fn()
            b.record()
        torch.cuda.synchronize()
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = _lib.launch_count() - n0
        if world > 1:
            tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax)
        return ms, launches

    ms_total, launches = timed(step, args.steps)
    for _ in range(max(args.warmup, 5)):        # the PCIe path (pinned staging copy, mapped result stores) needs its own warm-up
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None
    del flush
    torch.cuda.empty_cache()

    sg, streams = None, None
    if world > 1 and not args.no_scatter_gather:
        sg = scatter_gather(dec, dev, rank, world, max(2, min(args.steps, 5)), barrier)
        try:
            streams = streaming_share(dec, dev, rank, world, barrier)
        except Exception as e:              # the contract line must survive a failure of the extras (all ranks fail alike)
            streams = {"error": f"{type(e).__name__}: {e}"}

    # per-launcher event profile: a SEPARATE pass (graph replay off, two events per launch), reported as a breakdown
    # only -- every roofline figure below comes from the timed steps above
    prof = None
    if rank == 0:
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
        _lib.set_option("profile", "1")
        pk = 3
        for _ in range(pk):
            flush.fill_(1)
            step()
        prof = _lib.profile_report()
        _lib.set_option("profile", "0")
        del flush
        for v in prof.values():
            v["ms_per_step"] = v["ms"] / pk
            v["launches_per_step"] = v["launches"] / pk

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * samples / (ms_step * 1e-3)
        e2e_val = world * samples / (ms_e2e / args.steps * 1e-3)
        peak, peak_src = measured_peak_hbm()
        prof_ms = sum(v["ms_per_step"] for v in prof.values())
        conv_prof_ms = sum(v["ms_per_step"] for k, v in prof.items() if k.startswith(("conv1d", "tc_")))
        per_gpu_rate = samples / (ms_step * 1e-3)
        fp32_peak = _lib.measure_fp32_peak(dev)
        traffic = None
        tp = os.path.join(REPO, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("decoder_dram_bytes_per_step")
            except Exception:
                traffic = None
        achieved = CONV1D_BYTES_PER_SAMPLE * samples / (ms_step * 1e-3) / 1e9
        roofline = {
            "bound": "hbm",
            "kernel": "the decoder step as timed (one CUDA-graph replay of all its launches); Conv1d-layer traffic model",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_step": CONV1D_BYTES_PER_SAMPLE * samples,
            "kernel_ms_per_step": ms_step, "kernel_share_of_step": 1.0,
            "hbm_conv1d_fraction": per_gpu_rate * CONV1D_BYTES_PER_SAMPLE / (peak * 1e9),
            "hbm_compulsory_fraction": per_gpu_rate * COMPULSORY_BYTES_PER_SAMPLE / (peak * 1e9),
            "fp32_flop_fraction": per_gpu_rate * FLOP_PER_SAMPLE / (fp32_peak * 1e12),
            "fp32_peak_tflops_measured": fp32_peak, "fp32_peak_tflops_nominal": FP32_PEAK_TFLOPS_NOMINAL,
            "breakdown": {"note": "separate profiling pass (graphs off, events per launch: adds ~4 us per launch); shares only",
                          "dense_conv_share": conv_prof_ms / prof_ms if prof_ms else None,
                          "per_kernel_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items())}},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            utts = 16
            rate, ts = oracle_decoder_rate(sd_cpu, host, utts, reps=5, threads=threads)
            rate1, _ = oracle_decoder_rate(sd_cpu, host, 4, reps=2, threads=1)
            cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
                   "sample": f"{utts} of the {BATCH} utterances ({utts * LF * FRAME} samples), median of 5 after 1 warm-up, "
                             "oracle port of decoder.py on torch-CPU",
                   "value_1thread": rate1, "sample_1thread": f"4 utterances ({4 * LF * FRAME} samples), median of 2, 1 thread"}
        h2d = stage_host.numel() * 4                         # the one staged copy: content + f0 + energy (+ alignment gaps)
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random F0/content/weights)",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "frames": LF, "samples_per_step_per_gpu": samples,
                       "l2": "512 MB flush write before every timed step", "conv_impl": args.conv_impl,
                       "noise_draw": "in-kernel Philox per step (the reference's torch.rand, decoder.py:78)"},
            "e2e": {"value": e2e_val, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": out_pinned.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "rtf_x": value / 24000.0,
        }
        if sg is not None:
            line["scatter_gather"] = sg
            line["streams"] = streams
        if world == 1 and not args.no_extra_configs:
            try:
                line["other_configs"] = other_configs(dec, dev, peak)
            except Exception as e:          # the contract line must survive a failure of the extras
                line["other_configs"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    import faulthandler
    faulthandler.enable()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle timing (profiler runs)")
    ap.add_argument("--no-scatter-gather", action="store_true",
                    help="N > 1: skip the sharded configs[3] job (rank 0 holds 512 x N utterances of 10 s)")
    ap.add_argument("--no-extra-configs", action="store_true", help="N = 1: skip configs[2], the configs[3] share and the configs[4] share")
    ap.add_argument("--conv-impl", default=os.environ.get("TVC_CONV_IMPL", "tc"), choices=["fp32", "tc"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    from tinyvc_b200 import _lib
    _lib.set_option("conv_impl", args.conv_impl)
    run_ours(args)


if __name__ == "__main__":
    main()

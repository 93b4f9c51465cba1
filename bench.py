#!/usr/bin/env python
"""Contract benchmark: decoder audio samples/sec (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N = 1): BASELINE.json configs[1] -- Decoder-only forward, batch = 64, 8192-sample chunks
(autopadded to 18 frames = 8640 samples), seeded random F0 + content, random
weights of the reference's architecture.  N > 1: every rank runs the same per-GPU batch on its own
utterances (utterances are independent; no data-path collective) -> "scaling": "weak".

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
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # CUDA-core FMA peak at max clock (not measured)
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
    pinned = {k: v.pin_memory() for k, v in host.items()}
    out_pinned = torch.empty(BATCH, LF * FRAME).pin_memory()
    samples = BATCH * LF * FRAME
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    # The public call, as the reference's callers make it (infer.py:66 -> decoder.infer(content, f0, energy)): the
    # uniform noise draw of decoder.py:78 happens inside the step (in the noise kernel; torch.rand in the reference).
    dec.seed_noise(1234 + rank)

    def step():
        return dec.infer(inp["content"], inp["f0"], inp["energy"])

    e2e_dev = {k: torch.empty_like(inp[k]) for k in ("content", "f0", "energy")}    # fixed device staging buffers

    def step_e2e():
        for k in ("content", "f0", "energy"):
            e2e_dev[k].copy_(pinned[k], non_blocking=True)
        y = dec.infer(e2e_dev["content"], e2e_dev["f0"], e2e_dev["energy"])
        out_pinned.copy_(y, non_blocking=True)
        return y

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
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    clk = clocks.stop() if clocks else None

    # Optional (--scatter-gather, N > 1): the north-star's other multi-GPU mode -- rank 0 holds the whole job
    # (world x BATCH utterances), tinyvc_b200.shard scatters utterance blocks over NCCL, every rank converts, rank 0
    # gathers the waveforms.  Reported as an extra object; `value` stays the collective-free weak-scaling number.
    sg = None
    if world > 1 and args.scatter_gather:
        from tinyvc_b200.shard import ShardedDecoder
        sd = ShardedDecoder(dec, dev, micro_batch=32)
        full = None
        if rank == 0:
            full = {k: v.to(dev) for k, v in synth.decoder_inputs(BATCH * world, LF, seed=4321).items()}

        def step_sg():
            if rank == 0:
                return sd.infer(full["content"], full["f0"], full["energy"])
            return sd.infer()

        for _ in range(3):
            step_sg()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_sg()
        torch.cuda.synchronize()
        barrier()
        dt = time.perf_counter() - t0
        tmax = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        per = float(tmax) / args.steps
        in_b = sum(v.numel() * 4 for k, v in full.items() if k != "rand01") * (world - 1) // world if rank == 0 else 0
        sg = {"value": world * samples / per, "unit": "samples/s", "ms_per_step": per * 1e3, "utterances": BATCH * world,
              "scattered_bytes_per_step": in_b, "gathered_bytes_per_step": (world - 1) * samples * 4,
              "timing": "host clock around K steps between barriers, max over ranks (transfers are inside)"}

    # per-kernel event profile (separate pass; not part of the timed numbers)
    prof = None
    if rank == 0:
        _lib.set_option("profile", "1")
        pk = 3
        for _ in range(pk):
            flush.fill_(1)
            step()
        prof = _lib.profile_report()
        _lib.set_option("profile", "0")
        for v in prof.values():
            v["ms_per_step"] = v["ms"] / pk
            v["launches_per_step"] = v["launches"] / pk

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * samples / (ms_step * 1e-3)
        e2e_val = world * samples / (ms_e2e / args.steps * 1e-3)
        peak, peak_src = measured_peak_hbm()
        conv_ms = sum(v["ms_per_step"] for k, v in prof.items() if k.startswith(("conv1d", "tc_")))
        prof_ms = sum(v["ms_per_step"] for v in prof.values())
        per_gpu_rate = samples / (ms_step * 1e-3)
        traffic = None
        tp = os.path.join(REPO, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("conv1d_dram_bytes_per_step")
            except Exception:
                traffic = None
        achieved = CONV1D_BYTES_PER_SAMPLE * samples / (conv_ms * 1e-3) / 1e9 if conv_ms > 0 else None
        roofline = {
            "bound": "hbm", "kernel": "dense-conv launches of one step (all Conv1d layers: tc_conv_kernel / conv1d_f32_kernel)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
            "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_step": CONV1D_BYTES_PER_SAMPLE * samples,
            "kernel_ms_per_step": conv_ms, "kernel_share_of_step": conv_ms / prof_ms if prof_ms else None,
            "hbm_conv1d_fraction": per_gpu_rate * CONV1D_BYTES_PER_SAMPLE / (peak * 1e9),
            "hbm_compulsory_fraction": per_gpu_rate * COMPULSORY_BYTES_PER_SAMPLE / (peak * 1e9),
            "fp32_flop_fraction": per_gpu_rate * FLOP_PER_SAMPLE / (FP32_PEAK_TFLOPS * 1e12),
            "per_kernel_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(prof.items())},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            utts = 16
            rate, ts = oracle_decoder_rate(sd_cpu, host, utts, reps=5, threads=threads)
            cpu = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
                   "sample": f"{utts} of the {BATCH} utterances ({utts * LF * FRAME} samples), median of 5 after 1 warm-up, "
                             "oracle port of decoder.py on torch-CPU"}
        h2d = sum(pinned[k].numel() * 4 for k in ("content", "f0", "energy"))
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
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle timing (profiler runs)")
    ap.add_argument("--scatter-gather", action="store_true",
                    help="N > 1: also time rank-0-holds-the-job scatter -> convert -> gather over NCCL (extra JSON object)")
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

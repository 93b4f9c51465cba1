#!/usr/bin/env python
"""Real-time conversion entry point -- same command line as the reference's infer_streaming.py (:19-33).

Per 1920-sample block the reference reads int16 from PyAudio, scales by 1/32768, applies the input gain,
calls `StreamInfer.audio_callback`, applies the output gain and writes int16 (:83-97).  The audio device
I/O stays on the host (it is not part of the accelerated path); the window roll, Generator.convert and
the SOLA search/cross-fade run on the GPU with no host synchronisation other than the final copy of the
1920 output samples.  `pyaudio` is imported lazily so that everything but the device loop works without
it; `--dry-run N` feeds N blocks of silence instead of opening audio devices.
"""
import argparse
import sys

import torch


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="realtime inference")
    p.add_argument("-encp", "--encoder-path", default="./models/encoder.pt")
    p.add_argument("-decp", "--decoder-path", default="./models/decoder.pt")
    p.add_argument("-i", "--input", default=0, type=int)
    p.add_argument("-o", "--output", default=0, type=int)
    p.add_argument("-l", "--loopback", default=-1, type=int)
    p.add_argument("-idx", "--index", default="NONE")
    p.add_argument("-p", "--pitch-shift", default=0, type=float)
    p.add_argument("-t", "--target", default="target.wav")
    p.add_argument("-c", "--chunk", default=1920, type=int)
    p.add_argument("-e", "--extra", default=3840, type=int)
    p.add_argument("-d", "--device", default="cuda")
    p.add_argument("-sr", "--sample-rate", default=24000, type=int)
    p.add_argument("-ig", "--input-gain", default=0, type=float)
    p.add_argument("-og", "--output-gain", default=0, type=float)
    p.add_argument("-f0-est", "--f0-estimation", default="default", choices=["default", "fcpe", "dio", "harvest"])
    p.add_argument("--dry-run", default=0, type=int, help="process N silent blocks without audio devices (not in the reference)")
    return p


def db_gain(x: torch.Tensor, gain_db: float) -> torch.Tensor:
    """torchaudio.functional.gain: multiply by 10^(dB/20)."""
    return x if gain_db == 0 else x * (10.0 ** (gain_db / 20.0))


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    from infer import load_generator, load_target
    from module.infer import StreamInfer
    device = torch.device(args.device)
    generator = load_generator(args.encoder_path, args.decoder_path, device)
    stream_infer = StreamInfer(generator, pitch_shift=args.pitch_shift, block_size=args.chunk, device=device,
                               extra_size=args.extra, f0_estimation=args.f0_estimation)
    stream_infer.target = load_target(generator, args, device)
    stream_infer.init_buffer()

    def convert_block(pcm: torch.Tensor) -> torch.Tensor:
        x = db_gain(pcm.to(device, non_blocking=True) / 32768, args.input_gain)
        y = db_gain(stream_infer.audio_callback(x), args.output_gain)
        return (y * 32768).to(torch.int16).cpu()

    if args.dry_run:
        for _ in range(args.dry_run):
            convert_block(torch.zeros(args.chunk))
        print(f"dry run: {args.dry_run} blocks of {args.chunk} samples converted")
        return 0

    try:
        import numpy as np
        import pyaudio
    except ImportError as e:                                   # pragma: no cover - needs an audio stack
        raise SystemExit(f"infer_streaming.py needs pyaudio for device I/O ({e}); use --dry-run to test without it")
    audio = pyaudio.PyAudio()

    def open_stream(index: int, **kw):
        return audio.open(format=pyaudio.paInt16, rate=args.sample_rate, channels=1, **kw)

    s_in = open_stream(args.input, input_device_index=args.input, input=True)
    s_out = open_stream(args.output, output_device_index=args.output, output=True)
    s_loop = open_stream(args.loopback, output_device_index=args.loopback, output=True) if args.loopback != -1 else None
    print("Converting voice, Ctrl+C to stop conversion")
    while True:
        pcm = np.frombuffer(s_in.read(args.chunk), dtype=np.int16).astype(np.float32)
        data = convert_block(torch.from_numpy(pcm)).numpy().tobytes()
        s_out.write(data)
        if s_loop is not None:
            s_loop.write(data)


if __name__ == "__main__":
    sys.exit(main())

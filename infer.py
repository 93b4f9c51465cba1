#!/usr/bin/env python
"""File conversion entry point -- same command line as the reference's infer.py (infer.py:18-29):

    python infer.py -i ./inputs/ -o ./outputs/ -encp models/encoder.pt -decp models/decoder.pt \
                    [-idx models/index.pt | -t target.wav] [-p SEMITONES] -d cuda

The reference builds Encoder/Decoder/Generator (infer.py:33-38), makes the target either by encoding a
wav or by loading `index.pt` (:44-49), then converts every wav/ogg/mp3 under --inputs in one
`generator.convert` call per file (:60-69).  This script does the same through the CUDA path.
Differences, all deliberate:
  * the default device is `cuda` (there is no CPU path; `-d cpu` is refused with a clear error);
  * the waveform IS moved to the device (the reference forgets to, infer.py:62-66, which is why its
    script only works with `-d cpu`);
  * `-c/-b/-nc` are parsed and ignored exactly like the reference (infer.py:27-29,40-41 never use them);
  * `-f0-est` is accepted and ignored (Generator.convert never reads it, generator.py:26-34).
"""
import argparse
import glob
import os
import sys

import torch

SUPPORTED = ("wav", "ogg", "mp3")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="TinyVC file conversion on B200")
    p.add_argument("-i", "--inputs", default="./inputs/")
    p.add_argument("-o", "--outputs", default="./outputs/")
    p.add_argument("-encp", "--encoder-path", default="./models/encoder.pt")
    p.add_argument("-decp", "--decoder-path", default="./models/decoder.pt")
    p.add_argument("-f0-est", "--f0-estimation", default="default")
    p.add_argument("-idx", "--index", default="NONE")
    p.add_argument("-t", "--target", default="target.wav")
    p.add_argument("-d", "--device", default="cuda")
    p.add_argument("-p", "--pitch-shift", default=0.0, type=float)
    p.add_argument("-c", "--chunk-size", default=1920, type=int)
    p.add_argument("-b", "--buffer-size", default=4, type=int)
    p.add_argument("-nc", "--no-chunking", default=False, type=bool)
    return p


def load_generator(encoder_path: str, decoder_path: str, device: torch.device):
    from module.infer import Generator
    from module.tinyvc import Decoder, Encoder
    if device.type != "cuda":
        raise SystemExit(f"infer.py: device {device} is not CUDA; tinyvc_b200 has no CPU path (use -d cuda)")
    enc, dec = Encoder().eval(), Decoder().eval()
    enc.load_state_dict(torch.load(encoder_path, map_location="cpu"))
    dec.load_state_dict(torch.load(decoder_path, map_location="cpu"))
    return Generator(enc.to(device), dec.to(device))


def load_audio_24k(path: str, device: torch.device) -> torch.Tensor:
    """torchaudio.load decodes on the host; the resampling of infer.py:63-64 runs on the device (tvc_resample)."""
    import torchaudio
    from module.utils import resample
    wf, sr = torchaudio.load(path)
    return resample(wf.to(device), sr, 24000)


def load_target(generator, args, device: torch.device) -> torch.Tensor:
    if args.index == "NONE":
        tgt, _ = generator.encode(load_audio_24k(args.target, device))       # infer.py:45-47
        return tgt
    return torch.load(args.index, map_location="cpu").to(device)                 # [1,768,N]  (extract_index.py:58)


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    device = torch.device(args.device)
    generator = load_generator(args.encoder_path, args.decoder_path, device)
    tgt = load_target(generator, args, device)
    os.makedirs(args.outputs, exist_ok=True)
    paths = [p for fmt in SUPPORTED for p in glob.glob(os.path.join(args.inputs, "*." + fmt))]
    import torchaudio
    for path in paths:
        print(f"Converting {path} ...")
        wf = load_audio_24k(path, device).mean(dim=0, keepdim=True)
        out = generator.convert(wf, tgt, args.pitch_shift, args.f0_estimation, device).cpu()
        name = os.path.splitext(os.path.basename(path))[0]
        torchaudio.save(os.path.join(args.outputs, f"{name}.wav"), src=out, sample_rate=24000)
    return 0


if __name__ == "__main__":
    sys.exit(main())

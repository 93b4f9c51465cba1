#!/usr/bin/env python
"""Build `index.pt` -- same command line as the reference's extract_index.py (:14-20), on the GPU.

    python extract_index.py --dataset-cache dataset_cache -encp models/encoder.pt -size 2048 -o models/index.pt -d cuda

Output format is the reference's: `torch.save(tensor[1, 768, size] fp32)` (extract_index.py:58), which
`infer.py -idx` / `infer_streaming.py -idx` load.  See tinyvc_b200/index.py for how the clip selection
mirrors the reference's shuffled loader.
"""
import argparse
import sys

import torch


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="extract index")
    p.add_argument("--dataset-cache", default="dataset_cache")
    p.add_argument("-encp", "--encoder-path", default="models/encoder.pt")
    p.add_argument("-size", default=2048, type=int)
    p.add_argument("-o", "--output", default="models/index.pt")
    p.add_argument("-d", "--device", default="cuda")
    p.add_argument("--stride", default=4, type=int)
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    from module.tinyvc import Encoder
    from tinyvc_b200.index import build_index, cache_lengths, cache_loader
    device = torch.device(args.device)
    if device.type != "cuda":
        raise SystemExit(f"extract_index.py: device {device} is not CUDA; tinyvc_b200 has no CPU path (use -d cuda)")
    encoder = Encoder().eval()
    encoder.load_state_dict(torch.load(args.encoder_path, map_location="cpu"))
    encoder = encoder.to(device)
    print("Extracting...")
    tgt = build_index(encoder, cache_loader(args.dataset_cache), cache_lengths(args.dataset_cache), size=args.size,
                      stride=args.stride, device=device)
    print(f"Extracted {tgt.shape[2]} vectors")
    print("Saving...")
    torch.save(tgt, args.output)
    print("Complete")
    return 0


if __name__ == "__main__":
    sys.exit(main())

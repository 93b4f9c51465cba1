#!/usr/bin/env python
"""2..8-GPU check of tinyvc_b200.shard over NCCL (run with torchrun): rank 0 holds a decoder batch, ShardedDecoder
scatters utterance blocks, every rank converts, rank 0 gathers; the result must equal rank 0 converting the whole
batch alone, bit for bit.  Also times the sharded call (scatter + convert + gather) against the local one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_nccl_check.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from tinyvc_b200 import synth
from tinyvc_b200.shard import ShardedDecoder
from tinyvc_b200.tinyvc import Decoder
from tinyvc_b200.weights import load_synth_weights


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    dec = load_synth_weights(Decoder().eval(), seed=7).to(dev)
    B, Lf = int(os.environ.get("CHECK_B", 37)), int(os.environ.get("CHECK_LF", 50))      # ragged: 37 utterances
    sd = ShardedDecoder(dec, dev, micro_batch=int(os.environ.get("CHECK_MB", 8)), transport=os.environ.get("CHECK_TRANSPORT", "auto"),
                        direct_out=os.environ.get("CHECK_DIRECT", "1") == "1")
    inp = None
    if rank == 0:
        inp = {k: v.to(dev) for k, v in synth.decoder_inputs(B, Lf, seed=77).items()}
    res = {}
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        out = sd.infer(inp["content"], inp["f0"], inp["energy"], inp["rand01"]) if rank == 0 else sd.infer()
        torch.cuda.synchronize()
        dist.barrier()
        res["sharded_ms"] = (time.perf_counter() - t0) * 1e3
    if rank == 0:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ref = dec.infer(inp["content"], inp["f0"], inp["energy"], rand01=inp["rand01"])
        torch.cuda.synchronize()
        res["local_ms"] = (time.perf_counter() - t0) * 1e3
        res.update(transport=sd.transport, direct_out=sd.direct_out, world=world, utterances=B, frames=Lf, identical=bool(torch.equal(out, ref)),
                   max_abs_diff=float((out - ref).abs().max()))
        print(json.dumps(res), flush=True)
        assert res["identical"], "sharded result differs from the single-GPU result"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

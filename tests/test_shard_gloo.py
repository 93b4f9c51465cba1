"""Utterance sharding on CPU: world_size-2 and -3 `gloo` groups.

The scatter -> convert -> gather host logic is exercised with the CPU oracle standing in for the
per-rank conversion (the product path has no CPU arithmetic; `ShardedDecoder(decode=...)` exists for
exactly this).  Gate: the gathered result is bit-identical to converting the whole batch in one
process, for even, ragged and smaller-than-world batches.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from tinyvc_b200 import shard, synth  # noqa: E402


def test_partition_properties():
    for n in (0, 1, 2, 5, 64, 4096, 1023):
        for w in (1, 2, 3, 4, 8):
            blocks = shard.partition(n, w)
            assert len(blocks) == w
            assert sum(c for _, c in blocks) == n
            assert blocks[0][0] == 0
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s1 == s0 + c0                      # contiguous, ordered
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1
    assert shard.partition(4096, 8) == [(i * 512, 512) for i in range(8)]      # configs[3]: 512 utterances per GPU
    assert shard.micro_batches(5, 2) == [(0, 2), (2, 2), (4, 1)]
    assert shard.micro_batches(0, 2) == []
    assert [shard.stream_owner(s, 1024, 8) for s in (0, 127, 128, 1023)] == [0, 0, 1, 7]   # configs[4]: 128 streams/GPU
    with pytest.raises(ValueError):
        shard.partition(3, 0)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _toy_decode(content, f0, energy, rand01):
    """Cheap per-utterance function with the decoder's signature: each output row depends on every input
    of that utterance and on nothing else, so a misrouted or reordered shard changes the result."""
    B = content.shape[0]
    mix = content.reshape(B, -1).sum(1, keepdim=True) + f0.reshape(B, -1).sum(1, keepdim=True)
    if rand01 is not None:
        mix = mix + rand01.reshape(B, -1).mean(1, keepdim=True)
    return energy[:, 0, :] * 2.0 + mix


def _worker(rank, world, port, batch, lf, micro, use_oracle, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        decode = _toy_decode
        if use_oracle:
            import json
            from oracle import tinyvc_oracle as O
            from tinyvc_b200.weights import synth_state_dict
            keys = json.load(open(os.path.join(REPO, "tests", "golden", "state_keys.json")))["decoder"]
            sd = synth_state_dict({k: torch.empty(s) for k, s in keys}, 7)

            def decode(content, f0, energy, rand01):          # per utterance, so batching cannot change the bits
                with torch.inference_mode():
                    return torch.cat([O.decoder_infer(sd, content[i:i + 1], f0[i:i + 1], energy[i:i + 1], rand01[i:i + 1])
                                      for i in range(content.shape[0])])
        sd_ = shard.ShardedDecoder(None, torch.device("cpu"), micro_batch=micro, decode=decode)
        inp = synth.decoder_inputs(batch, lf, seed=99) if rank == 0 else None
        if rank == 0:
            got = sd_.infer(inp["content"], inp["f0"], inp["energy"], inp["rand01"])
            want = decode(inp["content"], inp["f0"], inp["energy"], inp["rand01"])
            q.put((tuple(got.shape), bool(torch.equal(got, want))))
        else:
            assert sd_.infer() is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(world, batch, lf, micro, use_oracle=False):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, lf, micro, use_oracle, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("sharded worker hung")
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    return q.get()


@pytest.mark.parametrize("world,batch,micro", [(2, 8, 2), (2, 5, 2), (3, 7, 64), (3, 2, 1), (2, 1, 4)])
def test_scatter_gather_matches_single_process(world, batch, micro):
    shape, same = _run(world, batch, lf=3, micro=micro)
    assert shape == (batch, 3 * 480)
    assert same, "gathered result differs from the single-process result"


def test_sharded_oracle_decoder_bit_identical():
    """Two ranks, the CPU oracle decoder per rank: utterance results do not depend on the rank count."""
    shape, same = _run(2, 3, lf=4, micro=1, use_oracle=True)
    assert shape == (3, 4 * 480) and same


def test_single_process_path_needs_no_group():
    inp = synth.decoder_inputs(5, 2, seed=3)
    sd_ = shard.ShardedDecoder(None, torch.device("cpu"), micro_batch=2, decode=_toy_decode)
    got = sd_.infer(inp["content"], inp["f0"], inp["energy"], inp["rand01"])
    assert torch.equal(got, _toy_decode(inp["content"], inp["f0"], inp["energy"], inp["rand01"]))

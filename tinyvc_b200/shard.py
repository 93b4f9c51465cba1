"""Utterance sharding across the GPUs of one box.

The reference has no distributed code.  Every utterance (or stream) on this path is independent
end to end -- GRN reduces over time and channels inside one item (reference convnext.py:32-33),
`match_features` loops per item (feature_retrieval.py:30) -- so the only multi-GPU structure is:

    rank 0 holds the batch  ->  scatter contiguous blocks of utterances  ->  every rank converts its
    block with the single-GPU path  ->  gather the waveforms on rank 0.

No reduction collective exists anywhere on the path.  One process per GPU (`torchrun`); the
exchange is grouped point-to-point over the process group's backend (NCCL over NVLink on the GPU
box, gloo on CPU for the host-logic tests), because blocks are ragged when B % world != 0.
Blocks are cut into micro-batches so that the transfer of micro-batch i+1 and the return of
micro-batch i-1 overlap the conversion of micro-batch i.

Per-utterance results do not depend on the rank count: a rank runs exactly the kernels the
single-GPU path runs on those utterances (no atomics, no cross-utterance reductions).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

Fields = Dict[str, torch.Tensor]


def partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks [(start, count)] of n utterances over `world` ranks; the first n % world
    ranks take one extra.  count may be 0 when n < world."""
    if n < 0 or world <= 0:
        raise ValueError(f"partition: bad arguments n={n} world={world}")
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        c = base + (1 if r < extra else 0)
        out.append((start, c))
        start += c
    return out


def micro_batches(count: int, size: int) -> List[Tuple[int, int]]:
    """[(offset, n)] covering `count` utterances in pieces of at most `size`."""
    if size <= 0:
        raise ValueError("micro-batch size must be positive")
    return [(o, min(size, count - o)) for o in range(0, count, size)]


@dataclass
class _Meta:
    """What a receiving rank must know before the payload arrives."""
    n: int                                  # utterances in the whole batch
    trailing: Dict[str, Tuple[int, ...]]    # per field: shape after the batch dimension
    order: Tuple[str, ...]


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _broadcast_meta(meta: Optional[_Meta], src: int, group) -> _Meta:
    box = [meta]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def sharded_map(fn: Callable[[Fields], torch.Tensor], fields: Optional[Fields], *, device: torch.device,
                out_trailing: Callable[[Fields], Tuple[int, ...]], micro_batch: int = 64, src: int = 0,
                group=None) -> Optional[torch.Tensor]:
    """Scatter -> fn -> gather.

    `fields` (rank `src` only; None elsewhere): name -> tensor [n, ...] on `device`, all with the same
    leading size n.  Every rank calls `fn` on its micro-batches (dict of tensors [m, ...]) and must return
    a tensor [m, *out_trailing(fields_of_that_micro_batch)] on `device`.  Returns the gathered result
    [n, ...] on rank `src`, None on the other ranks.  With no process group this is `fn(fields)` in
    micro-batches on the caller's device.
    """
    rank, world = _world(group)
    if world == 1:
        n = next(iter(fields.values())).shape[0]
        outs = [fn({k: v[o:o + m] for k, v in fields.items()}) for o, m in micro_batches(n, micro_batch)]
        if not outs:
            mb0 = {k: v[:0] for k, v in fields.items()}
            return torch.empty((0, *out_trailing(mb0)), device=device)
        return torch.cat(outs, dim=0)

    meta = None
    if rank == src:
        names = tuple(sorted(fields))
        sizes = {fields[k].shape[0] for k in names}
        if len(sizes) != 1:
            raise RuntimeError(f"sharded_map: fields disagree on the batch size: {sizes}")
        meta = _Meta(n=sizes.pop(), trailing={k: tuple(fields[k].shape[1:]) for k in names}, order=names)
    meta = _broadcast_meta(meta, src, group)
    blocks = partition(meta.n, world)
    gsrc = dist.get_global_rank(group, src) if group is not None else src

    def peer(r: int) -> int:
        return dist.get_global_rank(group, r) if group is not None else r

    # ---- exchange schedule: (rank, offset inside the batch, count), micro-batch major so that every rank's
    #      first micro-batch is on the wire before anyone's second
    per_rank = [micro_batches(c, micro_batch) for (_, c) in blocks]
    depth = max((len(p) for p in per_rank), default=0)
    my_start, my_count = blocks[rank]
    mine = per_rank[rank]

    # root: keep its own block as views, post sends of everyone else's micro-batches up front (async)
    send_work = []
    if rank == src:
        fields = {k: v.contiguous() for k, v in fields.items()}
        for j in range(depth):
            ops = []
            for r in range(world):
                if r == src or j >= len(per_rank[r]):
                    continue
                o, m = per_rank[r][j]
                g0 = blocks[r][0] + o
                for k in meta.order:
                    ops.append(dist.P2POp(dist.isend, fields[k][g0:g0 + m], peer(r), group))
            if ops:
                send_work.extend(dist.batch_isend_irecv(ops))
        result = None
        probe = {k: fields[k][:0] for k in meta.order}
        result = torch.empty((meta.n, *out_trailing(probe)), device=device, dtype=torch.float32)
        # post the receives of every remote micro-batch result
        recv_work = []
        for j in range(depth):
            ops = []
            for r in range(world):
                if r == src or j >= len(per_rank[r]):
                    continue
                o, m = per_rank[r][j]
                g0 = blocks[r][0] + o
                ops.append(dist.P2POp(dist.irecv, result[g0:g0 + m], peer(r), group))
            if ops:
                recv_work.extend(dist.batch_isend_irecv(ops))
        for o, m in mine:
            g0 = my_start + o
            result[g0:g0 + m] = fn({k: fields[k][g0:g0 + m] for k in meta.order})
        for w in send_work + recv_work:
            w.wait()
        return result

    # ---- non-root ranks: post every receive now, then convert micro-batch j as soon as it has landed and
    #      return it while j+1 is converted
    bufs: List[Fields] = []
    recv_work: List[list] = []
    for (o, m) in mine:
        mb = {k: torch.empty((m, *meta.trailing[k]), device=device, dtype=torch.float32) for k in meta.order}
        ops = [dist.P2POp(dist.irecv, mb[k], gsrc, group) for k in meta.order]
        recv_work.append(dist.batch_isend_irecv(ops))
        bufs.append(mb)
    back = []
    keep = []
    for j, mb in enumerate(bufs):
        for w in recv_work[j]:
            w.wait()
        y = fn(mb).contiguous()
        keep.append(y)
        back.extend(dist.batch_isend_irecv([dist.P2POp(dist.isend, y, gsrc, group)]))
    for w in back:
        w.wait()
    return None


class ShardedDecoder:
    """`Decoder.infer` over a batch held by rank 0, utterances sharded over the process group
    (BASELINE.json configs[3]: batch 4096 x 10 s over 8 GPUs).

        sd = ShardedDecoder(decoder, device)            # every rank, same weights
        wav = sd.infer(content, f0, energy, rand01)     # tensors on rank 0, None elsewhere -> [B, L] on rank 0
    """

    def __init__(self, decoder, device: torch.device, micro_batch: int = 64, group=None,
                 decode: Optional[Callable[..., torch.Tensor]] = None):
        self.decoder = decoder
        self.device = torch.device(device)
        self.micro_batch = micro_batch
        self.group = group
        # `decode` is injectable so the host logic can be exercised on CPU (gloo) in tests
        self._decode = decode or (lambda content, f0, energy, rand01: decoder.infer(content, f0, energy, rand01=rand01))

    def infer(self, content=None, f0=None, energy=None, rand01=None) -> Optional[torch.Tensor]:
        rank, _ = _world(self.group)
        fields = None
        if content is not None:
            fields = {"content": content, "f0": f0, "energy": energy}
            if rand01 is not None:
                fields["rand01"] = rand01

        def fn(mb: Fields) -> torch.Tensor:
            return self._decode(mb["content"], mb["f0"], mb["energy"], mb.get("rand01"))

        def out_trailing(mb: Fields) -> Tuple[int, ...]:
            return (mb["energy"].shape[-1],)

        return sharded_map(fn, fields, device=self.device, out_trailing=out_trailing, micro_batch=self.micro_batch,
                           group=self.group)


class ShardedGenerator:
    """`Generator.convert` over a batch of raw audio held by rank 0 (BASELINE.json configs[2] at N > 1):
    the scatter carries 4 B/sample of audio instead of the decoder's content features."""

    def __init__(self, generator, device: torch.device, micro_batch: int = 32, group=None,
                 convert: Optional[Callable[..., torch.Tensor]] = None):
        self.generator = generator
        self.device = torch.device(device)
        self.micro_batch = micro_batch
        self.group = group
        self._convert = convert or (lambda wf, tgt, shift, rand01: generator.convert(wf, tgt, shift, rand01=rand01))

    def convert(self, wf, tgt, pitch_shift: float = 0.0, rand01=None) -> Optional[torch.Tensor]:
        """wf [B, T] (T a multiple of 480) on rank 0, None elsewhere; `tgt` is replicated on every rank."""
        fields = None
        if wf is not None:
            if wf.shape[1] % 480:
                raise RuntimeError("ShardedGenerator.convert: pad the batch to a multiple of 480 samples first "
                                   "(utils.autopad_waveform)")
            fields = {"wf": wf}
            if rand01 is not None:
                fields["rand01"] = rand01

        def fn(mb: Fields) -> torch.Tensor:
            return self._convert(mb["wf"], tgt, pitch_shift, mb.get("rand01"))

        return sharded_map(fn, fields, device=self.device, out_trailing=lambda mb: (mb["wf"].shape[-1],),
                           micro_batch=self.micro_batch, group=self.group)


def stream_owner(stream_id: int, num_streams: int, world: int) -> int:
    """Rank that owns a stream for its lifetime (`input_wav` / `sola_buffer` state lives there;
    reference stream.py:58-65)."""
    for r, (s, c) in enumerate(partition(num_streams, world)):
        if s <= stream_id < s + c:
            return r
    raise IndexError(stream_id)


__all__: Sequence[str] = ("partition", "micro_batches", "sharded_map", "ShardedDecoder", "ShardedGenerator",
                          "stream_owner")

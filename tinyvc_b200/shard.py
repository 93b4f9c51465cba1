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


class _P2PState:
    """Per (buffer set) state of the one-sided transport: windows on the owner's tensors + staging buffers."""

    def __init__(self):
        self.key = None
        self.win: Dict[str, torch.Tensor] = {}
        self.n = 0
        self.names: Tuple[str, ...] = ()
        self.result: Optional[torch.Tensor] = None      # the gathered waveforms: owner's buffer, mapped on every rank
        self.window = None                              # peer.ResultWindow behind `result`
        self.stage: List[Fields] = []                   # two input staging sets (non-owners)
        self.ybuf: List[torch.Tensor] = []              # two output staging buffers (non-owners, copy-out mode)
        self.copy_in: Optional[torch.cuda.Stream] = None
        self.copy_out: Optional[torch.cuda.Stream] = None


class ShardedDecoder:
    """`Decoder.infer` over a batch held by rank 0, utterances sharded over the process group
    (BASELINE.json configs[3]: batch 4096 x 10 s over 8 GPUs).

        sd = ShardedDecoder(decoder, device)            # every rank, same weights
        wav = sd.infer(content, f0, energy, rand01)     # tensors on rank 0, None elsewhere -> [B, L] on rank 0

    transport:
      "p2p"  (default on CUDA when every GPU can address rank 0's): one-sided.  Rank 0 publishes its input and result
             tensors as peer windows (tinyvc_b200.peer, CUDA IPC); every other rank pulls its utterance block out of
             them micro-batch by micro-batch with copy-engine peer copies on a side stream (the pull of micro-batch
             j + 1 overlaps the conversion of j) and the LAST KERNEL of the conversion stores the waveform straight
             into rank 0's result tensor over NVLink (`Decoder.infer(out=window slice)`), so there is no gather step.
             The process group carries control only: window handles once per buffer set, two barriers per call.
      "nccl" two-sided grouped send/recv through the process group's backend (`sharded_map`); the only choice on CPU
             (gloo, tests) and the fallback when peer access is unavailable.
    Per-utterance results are those of the single-GPU path either way (same kernels on the same utterances).
    """

    def __init__(self, decoder, device: torch.device, micro_batch: int = 64, group=None,
                 decode: Optional[Callable[..., torch.Tensor]] = None, transport: str = "auto", direct_out: bool = True):
        self.decoder = decoder
        self.device = torch.device(device)
        self.micro_batch = micro_batch
        self.group = group
        self.direct_out = direct_out
        # `decode` is injectable so the host logic can be exercised on CPU (gloo) in tests
        self._custom = decode is not None
        self._decode = decode or (lambda content, f0, energy, rand01, out=None:
                                  decoder.infer(content, f0, energy, rand01=rand01, out=out))
        rank, world = _world(self.group)
        if transport == "auto":
            transport = "p2p" if (self.device.type == "cuda" and world > 1 and not self._custom) else "nccl"
        if transport == "p2p" and world > 1:
            # all ranks must agree: p2p only if every rank's GPU can address the owner's GPU
            from . import peer
            owner_dev = torch.tensor([self.device.index if rank == 0 else -1], device=self.device)
            dist.all_reduce(owner_dev, op=dist.ReduceOp.MAX, group=self.group)
            ok = torch.tensor([1 if peer.p2p_available(self.device, int(owner_dev)) else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok) == 0:
                transport = "nccl"
        self.transport = transport
        self._p2p = _P2PState()
        if decoder is not None and self.device.type == "cuda" and world > 1 and hasattr(decoder, "seed_noise"):
            # in-kernel noise draws (no injected rand01): distinct streams per rank, or utterance u of every rank's
            # micro-batch j would get the same draw
            decoder.seed_noise(0x5EED0000 + 7919 * rank)

    # ---- one-sided transport --------------------------------------------------------------------------------
    def _bind(self, fields: Optional[Fields]) -> None:
        """Collective.  (Re)publishes rank 0's tensors as peer windows when the buffer set changed."""
        from . import peer
        rank, world = _world(self.group)
        st = self._p2p
        hdr = torch.zeros(4, dtype=torch.int64, device=self.device)
        if rank == 0:
            names = tuple(sorted(fields))
            key = tuple((k, fields[k].data_ptr(), tuple(fields[k].shape)) for k in names)
            changed = key != st.key
            n = fields[names[0]].shape[0]
            hdr[0], hdr[1] = int(changed), n
        dist.broadcast(hdr, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        if int(hdr[0]) == 0:
            return
        if rank == 0:
            for k in names:
                if not fields[k].is_contiguous():
                    raise RuntimeError(f"ShardedDecoder: {k} must be contiguous on rank 0 (it is published as a peer window)")
            st.key, st.names, st.n = key, names, n
            st.win = peer.share_from(0, {k: fields[k] for k in names}, self.group)
            L = int(fields["energy"].shape[-1])
            hdr2 = torch.tensor([L], dtype=torch.int64, device=self.device)
        else:
            st.win = peer.share_from(0, None, self.group)
            st.names = tuple(sorted(st.win))
            st.n = int(hdr[1])
            st.key = ("remote", st.n)
            st.stage, st.ybuf = [], []
            hdr2 = torch.zeros(1, dtype=torch.int64, device=self.device)
            # first touch of the owner's GPU from this process: a tiny peer copy makes torch set the copy path up
            probe = torch.empty(1, device=self.device)
            probe.copy_(st.win[st.names[0]].view(-1)[:1])
            torch.cuda.synchronize(self.device)
        dist.broadcast(hdr2, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        # the result lives in a buffer the library allocated on rank 0 and every rank mapped under its own device, so that
        # the conversion's last kernel can store into it from any GPU (peer.ResultWindow)
        if st.window is not None:
            st.window.close()
        st.window = peer.ResultWindow(0, (st.n, int(hdr2[0])), self.device, self.group)
        st.result = st.window.tensor
        if st.copy_in is None:
            st.copy_in = torch.cuda.Stream(self.device)
            st.copy_out = torch.cuda.Stream(self.device)

    def close(self) -> None:
        """Collective.  Releases the peer windows: the other ranks drop their mappings of rank 0's tensors first, then rank 0
        frees the result buffer.  (The tensor `infer` returned on rank 0 aliases that buffer: clone it to keep it.)"""
        rank, world = _world(self.group)
        st = self._p2p
        if world > 1 and self.transport == "p2p" and st.key is not None:
            torch.cuda.synchronize(self.device)
            if rank != 0:
                st.win, st.stage, st.ybuf = {}, [], []
                if st.window is not None:
                    st.window.close()
                    st.window = None
                import gc
                gc.collect()
                torch.cuda.ipc_collect()               # torch releases its IPC references lazily
            dist.barrier(group=self.group)
            if rank == 0:
                st.win = {}
                if st.window is not None:
                    st.window.close()
                    st.window = None
            st.result = None
            st.key = None
            dist.barrier(group=self.group)

    def _infer_p2p(self, fields: Optional[Fields]) -> Optional[torch.Tensor]:
        rank, world = _world(self.group)
        self._bind(fields)
        st = self._p2p
        start, count = partition(st.n, world)[rank]
        mbs = micro_batches(count, self.micro_batch)
        cur = torch.cuda.current_stream(self.device)
        if rank == 0:
            cur.synchronize()                      # the inputs other ranks are about to pull are complete
        dist.barrier(group=self.group)
        if rank == 0:
            for o, m in mbs:
                g0 = start + o
                self._decode(*(fields[k][g0:g0 + m] if k in fields else None for k in ("content", "f0", "energy", "rand01")),
                             out=st.result[g0:g0 + m])
            cur.synchronize()
            dist.barrier(group=self.group)         # every rank's stores / copies into `result` have completed
            return st.result
        names = st.names
        if not st.stage:
            mmax = min(self.micro_batch, max(count, 1))
            st.stage = [{k: torch.empty((mmax, *st.win[k].shape[1:]), device=self.device, dtype=torch.float32) for k in names}
                        for _ in range(2)]
            if not self.direct_out:
                st.ybuf = [torch.empty((mmax, st.result.shape[1]), device=self.device, dtype=torch.float32) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in mbs]
        done = [torch.cuda.Event() for _ in mbs]
        pushed = [torch.cuda.Event() for _ in mbs]

        def pull(j):
            o, m = mbs[j]
            g0 = start + o
            with torch.cuda.stream(st.copy_in):
                if j >= 2:
                    st.copy_in.wait_event(done[j - 2])            # the staging set is free again
                for k in names:
                    st.stage[j & 1][k][:m].copy_(st.win[k][g0:g0 + m], non_blocking=True)
                ready[j].record(st.copy_in)

        for j in range(min(2, len(mbs))):
            pull(j)
        for j, (o, m) in enumerate(mbs):
            g0 = start + o
            cur.wait_event(ready[j])
            mb = {k: st.stage[j & 1][k][:m] for k in names}
            if self.direct_out:
                self._decode(mb["content"], mb["f0"], mb["energy"], mb.get("rand01"), out=st.result[g0:g0 + m])
                done[j].record(cur)
            else:
                if j >= 2:
                    cur.wait_event(pushed[j - 2])
                y = st.ybuf[j & 1][:m]
                self._decode(mb["content"], mb["f0"], mb["energy"], mb.get("rand01"), out=y)
                done[j].record(cur)
                with torch.cuda.stream(st.copy_out):
                    st.copy_out.wait_event(done[j])
                    st.result[g0:g0 + m].copy_(y, non_blocking=True)
                    pushed[j].record(st.copy_out)
            if j + 2 < len(mbs):
                pull(j + 2)
        cur.synchronize()
        st.copy_out.synchronize()
        dist.barrier(group=self.group)
        return None

    def infer(self, content=None, f0=None, energy=None, rand01=None) -> Optional[torch.Tensor]:
        rank, world = _world(self.group)
        fields = None
        if content is not None:
            fields = {"content": content, "f0": f0, "energy": energy}
            if rand01 is not None:
                fields["rand01"] = rand01
        if self.transport == "p2p" and world > 1:
            return self._infer_p2p(fields)

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

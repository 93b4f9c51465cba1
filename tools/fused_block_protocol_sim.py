"""Randomised discrete-event model of the fused Upsample block's synchronisation protocol (csrc/tc_block.cu).

Agents: the TMA producer thread, the MMA-issuing warp (tcgen05 ops retire asynchronously, in order; tcgen05.commit
arrives on a barrier when everything issued before it has retired), 12 epilogue warps.  mbarriers follow the PTX
semantics that matter here: a phase completes when the pending arrivals and the expected bytes reach zero, and
`try_wait.parity(P)` succeeds iff the number of completed phases is odd/even opposite to P -- so a waiter that is two
phases late blocks forever and one that asks two phases early falls through (both show up here as a deadlock or a
version mismatch).  Every buffer region carries a version tag; reads are checked when an op is issued AND when it
retires (a write in between = WAR hazard), TMA writes are checked against in-flight readers.

    python tools/fused_block_protocol_sim.py [runs] [segments]
"""
import os
import random
import sys

BUG = os.environ.get("SIM_BUG", "")      # self-test: "naive_wait" | "no_in_free" | "one_barrier" must each be caught

BT, RING, EPI = 4, 2, 12


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.done = count, count, 0, 0

    def _check(self):
        if self.pending == 0 and self.tx == 0:
            self.done += 1
            self.pending = self.count

    def arrive(self, tx=0):
        self.tx += tx
        self.pending -= 1
        assert self.pending >= 0, "too many arrivals"
        self._check()

    def complete_tx(self, n):
        self.tx -= n
        self._check()

    def test(self, parity):
        return (self.done & 1) != parity


class Sim:
    def __init__(self, nseg, seed):
        self.rng = random.Random(seed)
        self.nseg = nseg
        self.now = 0
        self.events = []                       # (time, seq, fn)
        self.seq = 0
        self.tc_free_at = 0                    # tensor-core queue: ops retire in order
        self.wfull, self.in_full, self.in_free = Bar(1), Bar(1), Bar(1)
        self.cond_full = [Bar(1) for _ in range(RING)]
        self.cond_empty = [Bar(1) for _ in range(RING)]
        self.acc_full = [Bar(1) for _ in range(2)]
        self.acc_empty = [Bar(EPI) for _ in range(2)]
        self.act_ready = [Bar(EPI) for _ in range(BT)]
        # version tags
        self.buf = [[[None] * EPI for _ in range(BT)] for _ in range(2)]     # [buffer][tile][epilogue warp part]
        self.buf_tma_inflight = [0, 0]
        self.buf_readers = [0, 0]              # MMA ops in flight that read the buffer
        self.cond = [None] * RING
        self.cond_readers = [0] * RING
        self.cond_tma_inflight = [0] * RING
        self.acc = [None, None]
        self.acc_reads_left = [0, 0]
        self.errors = []

    def at(self, delay, fn):
        self.seq += 1
        self.events.append((self.now + delay, self.seq, fn))

    def err(self, msg):
        self.errors.append(f"t={self.now}: {msg}")

    # ---- tensor-core queue
    def tc_issue(self, on_retire):
        t = max(self.tc_free_at, self.now) + self.rng.randint(1, 12)
        self.tc_free_at = t
        self.at(t - self.now, on_retire)

    def tc_commit(self, bar):
        t = max(self.tc_free_at, self.now)
        self.at(t - self.now + 1, bar.arrive)

    # ---- agents (generators yield ("wait", bar, parity) or ("step",))
    def producer(self):
        self.wfull.arrive(tx=1)
        self.at(self.rng.randint(5, 60), lambda: self.wfull.complete_tx(1))
        cs, cph, cwrapped = 0, 0, False
        for it in range(self.nseg):
            b = it & 1
            if it > 0 and BUG != "no_in_free":
                yield ("wait", self.in_free, (it - 1) & 1)
            if self.buf_readers[b]:
                self.err(f"TMA into buffer {b} for segment {it} while {self.buf_readers[b]} MMA ops still read it")
            self.in_full.arrive(tx=2)
            self.buf_tma_inflight[b] += 2
            for _ in range(2):
                def land(b=b, it=it):
                    if self.buf_readers[b]:
                        self.err(f"TMA landed in buffer {b} (segment {it}) under in-flight MMA reads")
                    self.buf_tma_inflight[b] -= 1
                    if self.buf_tma_inflight[b] == 0:
                        for j in range(BT):
                            self.buf[b][j] = [(it, -1)] * EPI
                    self.in_full.complete_tx(1)
                self.at(self.rng.randint(5, 80), land)
            yield ("step",)
            for pas in range(2):
                for j in range(BT):
                    if cwrapped:
                        yield ("wait", self.cond_empty[cs], cph)
                    if self.cond_readers[cs]:
                        self.err(f"cond stage {cs} refilled under in-flight reads")
                    self.cond_full[cs].arrive(tx=1)
                    self.cond_tma_inflight[cs] += 1

                    def cland(cs=cs, tag=(it, pas, j)):
                        if self.cond_readers[cs]:
                            self.err(f"cond stage {cs} landed under in-flight reads")
                        self.cond[cs] = tag
                        self.cond_tma_inflight[cs] -= 1
                        self.cond_full[cs].complete_tx(1)
                    self.at(self.rng.randint(5, 60), cland)
                    cs += 1
                    if cs == RING:
                        cs, cph, cwrapped = 0, (cph ^ 1 if cwrapped else 0), True
                    yield ("step",)

    def mma(self):
        yield ("wait", self.wfull, 0)
        tcount, cs, cph = 0, 0, 0
        seen = [0] * BT
        for it in range(self.nseg):
            bin_, bot = it & 1, (it & 1) ^ 1
            yield ("wait", self.in_full, it & 1)
            for l in range(5):
                src = bot if (l & 1) else bin_
                for j in range(BT):
                    buf, buse = tcount & 1, tcount >> 1
                    if buse > 0:
                        yield ("wait", self.acc_empty[buf], (buse - 1) & 1)
                    nbr = [jj for jj in range(BT) if j - 1 <= jj <= j + 1]
                    if l > 0:
                        need = it * 4 + l
                        if BUG == "one_barrier":
                            nbr_w = [j]                      # waits for its own tile only
                        else:
                            nbr_w = nbr
                        for jj in nbr_w:
                            if BUG == "naive_wait":          # re-waits a phase it has already seen pass
                                yield ("wait", self.act_ready[jj], (need - 1) & 1)
                                continue
                            while seen[jj] < need:
                                yield ("wait", self.act_ready[jj], seen[jj] & 1)
                                seen[jj] += 1
                    if self.acc_reads_left[buf]:
                        self.err(f"accumulator {buf} overwritten with {self.acc_reads_left[buf]} epilogue reads outstanding")
                    want = (it, l - 1)
                    reads = nbr if l < 4 else [j]

                    def check(when, src=src, reads=reads, want=want, it=it, l=l, j=j):
                        if self.buf_tma_inflight[src]:
                            self.err(f"MMA seg {it} layer {l} tile {j} {when}: buffer {src} has a TMA write in flight")
                        for jj in reads:
                            if any(tag != want for tag in self.buf[src][jj]):
                                self.err(f"MMA seg {it} layer {l} tile {j} {when}: buffer {src} tile {jj} holds {set(self.buf[src][jj])}, wants {want}")
                    check("issue")
                    self.buf_readers[src] += 1

                    def retire(src=src, buf=buf, tag=(it, l, j), check=check):
                        check("retire")
                        self.buf_readers[src] -= 1
                        self.acc[buf] = tag
                    self.tc_issue(retire)
                    if l in (1, 3):
                        yield ("wait", self.cond_full[cs], cph)
                        want_c = (it, l >> 1, j)
                        if self.cond[cs] != want_c or self.cond_tma_inflight[cs]:
                            self.err(f"aux MMA seg {it} layer {l} tile {j}: cond stage {cs} holds {self.cond[cs]}, wants {want_c}")
                        self.cond_readers[cs] += 1

                        def cretire(cs=cs, want_c=want_c):
                            if self.cond[cs] != want_c:
                                self.err(f"cond stage {cs} changed under an aux MMA ({self.cond[cs]} vs {want_c})")
                            self.cond_readers[cs] -= 1
                        self.tc_issue(cretire)
                        self.tc_commit(self.cond_empty[cs])
                        cs += 1
                        if cs == RING:
                            cs, cph = 0, cph ^ 1
                    self.acc_reads_left[buf] = EPI       # set when the commit fires would be later; conservative: now
                    self.tc_commit(self.acc_full[buf])
                    tcount += 1
                    yield ("step",)
                if l == 3:
                    self.tc_commit(self.in_free)

    def epilogue(self, w):
        tcount = 0
        for it in range(self.nseg):
            bin_, bot = it & 1, (it & 1) ^ 1
            for l in range(5):
                out = bin_ if (l & 1) else bot
                for j in range(BT):
                    buf, buse = tcount & 1, tcount >> 1
                    yield ("wait", self.acc_full[buf], buse & 1)
                    if self.acc[buf] != (it, l, j):
                        self.err(f"epilogue warp {w} seg {it} layer {l} tile {j}: accumulator {buf} holds {self.acc[buf]}")
                    for _ in range(self.rng.randint(0, 6)):
                        yield ("step",)
                    if l < 4:
                        if self.buf_readers[out] and False:
                            pass
                        if self.buf_tma_inflight[out]:
                            self.err(f"epilogue writes buffer {out} (seg {it} layer {l}) under a TMA write")
                        self.buf[out][j][w] = (it, l)
                    self.acc_reads_left[buf] -= 1
                    self.acc_empty[buf].arrive()
                    if l < 4:
                        self.act_ready[j].arrive()
                    tcount += 1

    def run(self):
        agents = {"prod": self.producer(), "mma": self.mma()}
        for w in range(EPI):
            agents[f"epi{w}"] = self.epilogue(w)
        waiting = {k: None for k in agents}
        while agents:
            self.now += 1
            due = sorted(e for e in self.events if e[0] <= self.now)
            self.events = [e for e in self.events if e[0] > self.now]
            for _, _, fn in due:
                fn()
            runnable = [k for k in agents if waiting[k] is None or waiting[k][0].test(waiting[k][1])]
            if not runnable:
                if not self.events:
                    self.err("DEADLOCK: " + ", ".join(f"{k} waits parity {v[1]} on a barrier with {v[0].done} phases done" for k, v in waiting.items() if v))
                    return
                continue
            k = self.rng.choice(runnable)
            waiting[k] = None
            try:
                r = next(agents[k])
            except StopIteration:
                del agents[k]
                del waiting[k]
                continue
            if r[0] == "wait":
                waiting[k] = (r[1], r[2])
            if self.errors:
                return


def main() -> int:
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    nseg = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    bad = 0
    for seed in range(runs):
        s = Sim(nseg, seed)
        s.run()
        if s.errors:
            bad += 1
            print(f"seed {seed}:", *s.errors[:4], sep="\n  ")
            if bad >= 3:
                break
    print(f"{runs} randomised schedules x {nseg} segments: {'no protocol violations' if not bad else f'{bad} FAILED'}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

"""CPU model of the fused Upsample block's window / slot / replicate-padding logic (csrc/tc_block.cu).

It does not model the hardware (TMA, mbarriers, tcgen05); it replays the kernel's *indexing*: which buffer slot every
MMA tap reads, which rows every epilogue writes, the replicate slots, the stale contents buffers keep from earlier
windows, the rows a window is allowed to emit -- and compares the emitted rows with a plain layer-by-layer
replicate-padded evaluation of the block.  Run on the CPU:  python tools/fused_block_model.py
"""
import sys

import numpy as np

BM, BT = 128, 4
BW = BM * BT
HALO = 40
BS = BW - 2 * HALO
PAD = 32
SLOTS = BW + 72
DIL = (1, 3, 9, 27)
C = 3


def lrelu(v):
    return np.where(v > 0, v, 0.1 * v)


def conv3(x, w, d):
    """x [T, C] replicate-padded k=3 dilated conv, w [3, C, C]"""
    T = x.shape[0]
    idx = np.arange(T)
    out = np.zeros_like(x)
    for tap in range(3):
        src = np.clip(idx + (tap - 1) * d, 0, T - 1)
        out += x[src] @ w[tap]
    return out


def reference(p, xi, cond, W):
    h1 = lrelu(conv3(p, W["c1"], DIL[0]))
    y = conv3(h1, W["c2"], DIL[1]) * (cond @ W["f1s"]) + cond @ W["f1h"] + xi
    h3 = lrelu(conv3(lrelu(y), W["c3"], DIL[2]))
    z = conv3(h3, W["c4"], DIL[3]) * (cond @ W["f2s"]) + cond @ W["f2h"] + y
    return z @ W["c5"]


def fused(p_all, xi_all, cond_all, W, B, T, rng):
    rows = B * T
    xo = np.full((rows, C), np.nan)
    written = np.zeros(rows, dtype=int)
    spu = -(-T // BS)
    # both buffers start as zeros; afterwards they keep whatever earlier windows left (finite garbage)
    bufs = [np.zeros((SLOTS, C)), np.zeros((SLOTS, C))]
    it = 0
    order = list(range(B * spu))
    for seg in order:                     # one CTA walking all segments is the worst case for stale data
        bq, k = divmod(seg, spu)
        baseT = bq * T
        w0 = k * BS - HALO
        R0 = (baseT + w0 - PAD) & ~7
        sh = baseT + w0 - R0
        assert 32 <= sh < 40
        bin_, bot = bufs[it & 1], bufs[(it & 1) ^ 1]
        # TMA: slot s <- operand row R0 + s, zeros outside the tensor (neighbouring utterances' rows arrive as they are)
        for s in range(SLOTS):
            R = R0 + s
            bin_[s] = p_all[R] if 0 <= R < rows else 0.0
        # input replicate fix-up (c1 has dilation 1)
        if w0 <= 0:
            s0 = sh - w0
            bin_[s0 - 1] = bin_[s0]
        if T - 1 < w0 + BW:
            sT = sh + (T - 1 - w0)
            bin_[sT + 1] = bin_[sT]
        y_tmem = np.zeros((BW, C))

        def mma(src, w, d, taps=3):
            acc = np.zeros((BW, C))
            for j in range(BT):
                for tap in range(taps):
                    start = sh + BM * j - d + tap * d if taps == 3 else sh + BM * j
                    assert 0 <= start and start + BM <= SLOTS, (start, d)
                    acc[j * BM:(j + 1) * BM] += src[start:start + BM] @ (w[tap] if taps == 3 else w)
            return acc

        def cond_tile():
            c = np.zeros((BW, C))
            for r in range(BW):
                R = baseT + w0 + r
                c[r] = cond_all[R] if 0 <= R < rows else 0.0
            return c

        def publish(dst, v, nd):
            for r in range(BW):
                t = w0 + r
                if not (0 <= t < T):
                    continue
                sl = sh + r
                dst[sl] = v[r]
                if t == 0:
                    for q in range(1, nd + 1):
                        dst[sl - q] = v[r]
                if t == T - 1:
                    for q in range(1, nd + 1):
                        dst[sl + q] = v[r]

        xi_w = np.zeros((BW, C))
        for r in range(BW):
            t = w0 + r
            if 0 <= t < T:
                xi_w[r] = xi_all[baseT + t]
        cnd = cond_tile()
        publish(bot, lrelu(mma(bin_, W["c1"], DIL[0])), DIL[1])                       # L1: in -> other
        y = mma(bot, W["c2"], DIL[1]) * (cnd @ W["f1s"]) + cnd @ W["f1h"] + xi_w       # L2: other -> in
        y_tmem[:] = y
        publish(bin_, lrelu(y), DIL[2])
        publish(bot, lrelu(mma(bin_, W["c3"], DIL[2])), DIL[3])                       # L3: in -> other
        z = mma(bot, W["c4"], DIL[3]) * (cnd @ W["f2s"]) + cnd @ W["f2h"] + y_tmem     # L4: other -> in
        publish(bin_, z, 0)
        o = mma(bin_, W["c5"], 0, taps=1)                                             # L5
        for r in range(HALO, HALO + BS):
            t = w0 + r
            if 0 <= t < T:
                xo[baseT + t] = o[r]
                written[baseT + t] += 1
        it += 1
    return xo, written


def main() -> int:
    rng = np.random.default_rng(0)
    bad = 0
    for B, T in ((1, 480), (2, 960), (3, 431), (2, 432), (2, 433), (1, 1300), (2, 8640 // 4), (1, 30), (2, 1)):
        rows = B * T
        W = {k: rng.normal(size=(3, C, C)) * 0.4 for k in ("c1", "c2", "c3", "c4")}
        for k in ("f1s", "f1h", "f2s", "f2h", "c5"):
            W[k] = rng.normal(size=(C, C)) * 0.5
        p, xi, cond = (rng.normal(size=(rows, C)) for _ in range(3))
        ref = np.concatenate([reference(p[b * T:(b + 1) * T], xi[b * T:(b + 1) * T], cond[b * T:(b + 1) * T], W) for b in range(B)])
        got, written = fused(p, xi, cond, W, B, T, rng)
        ok = np.all(written == 1) and np.allclose(got, ref, rtol=1e-10, atol=1e-10)
        print(f"B={B} T={T}: rows written once={bool(np.all(written == 1))} max|d|={np.nanmax(np.abs(got - ref)):.2e} {'ok' if ok else 'MISMATCH'}")
        bad += 0 if ok else 1
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

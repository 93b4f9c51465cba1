// Fused Upsample block at the FilterNet's highest rate (24 channels), sm_100a.  EXPERIMENTAL (default off).
//
//   p  = lrelu(x)                       (input planes, produced by the resampler)
//   h1 = lrelu(c1(p))                   k = 3, dil 1
//   y  = FiLM1(c2(h1); cond) + x        k = 3, dil 3          (module/tinyvc/decoder.py:165-171, FiLM :88-97)
//   h3 = lrelu(c3(lrelu(y)))            k = 3, dil 9
//   z  = FiLM2(c4(h3); cond) + y        k = 3, dil 27
//   xo = c5(z)                          1 x 1
//
// Unfused, the five convs move ~1.5 KB per time row through L2 / HBM (every 24-channel intermediate is written as
// split planes and read back, the fp32 residual too) and are bandwidth-bound.  Here a CTA owns a WINDOW of 512 rows of
// one utterance (4 MMA tiles of 128 rows), keeps every intermediate in shared memory as ready-made UMMA operands
// (chunk-major split planes are exactly the K-major smem order) and the fp32 residual `y` in TMEM, and produces the
// 432 rows in the middle of the window: the 40 rows on either side (1 + 3 + 9 + 27) are recomputed by the
// neighbouring windows.  Per row it reads p, x and cond (cond twice, the second time from L2) and writes xo.
//
// The arithmetic (MMA order per tile, epilogue operations) is that of tc_conv.cu, so the result is expected to be
// bit-identical to the five separate launches; tools/fused_block_check.py compares the two.
//
// Shared memory (bytes):   [buffer A: hi|lo planes, 3 chunks x 584 slots x 16 B][buffer B: same][cond ring: 2 x (hi|lo, 3 chunks
//   x 136 slots)][weight images of c1..c5, resident][mbarriers].  Slot s of a buffer holds operand row R0 + s (R0 a
//   multiple of 8 so that one TMA box per plane lands the input in place); window row 0 sits at slot sh in [32, 40).
//   Activations have 24 channels but a K-step is 16: the second K-step's upper chunk aliases the first chunk of
//   whatever follows (lo plane, next buffer, cond ring, weights -- all finite bf16, all zero-initialised) and meets
//   zero weights.
// Warp roles (448 threads): warps 0-11 epilogue (4 TMEM lane quarters x 3 column groups), warp 12 MMA issue,
//   warp 13 TMA producer (weights once, one window of p per segment, a cond tile per FiLM tile).
// Layer l + 1 of tile j starts once the epilogues of layer l for tiles j - 1, j, j + 1 have published their rows
// (act_ready[j] mbarriers); buffers ping-pong (L1: in -> other, L2: other -> in, ...); the next segment's input is
// fetched into `other` as soon as c4's MMAs have retired (in_free), so it lands behind c5 and the epilogues.
#include <cuda.h>
#include <cstdlib>
#include <cstring>

#include "tc_block.cuh"
#include "tc_ptx.cuh"

namespace tvc {

namespace {

constexpr int kBM = 128;                      // rows per MMA tile
constexpr int kBT = 4;                        // MMA tiles per window
constexpr int kBW = kBM * kBT;                // window rows
constexpr int kBHalo = 40;                    // 1 + 3 + 9 + 27: rows a window cannot produce on either side
constexpr int kBS = kBW - 2 * kBHalo;         // rows a window produces (432)
constexpr int kBPad = 32;                     // slots in front of window row 0 (>= the largest dilation, multiple of 8)
constexpr int kBSlots = kBW + 72;             // pad + alignment slack (< 8) + window + 27, rounded up to a multiple of 8
constexpr int kBMaxDil = 27;
constexpr uint32_t kActLbo = kBSlots * 16;    // bytes between the 8-channel chunk columns of an activation buffer
constexpr uint32_t kActPlane = 3 * kActLbo;   // hi -> lo plane
constexpr uint32_t kActBuf = 2 * kActPlane;
constexpr int kCondG = 17;                    // row groups of a cond tile (128 rows + up to 7 leading rows)
constexpr uint32_t kCondLbo = kCondG * 128;
constexpr uint32_t kCondPlane = 3 * kCondLbo;
constexpr uint32_t kCondStage = 2 * kCondPlane;
constexpr uint32_t kCondRing = 2;
constexpr uint32_t kOffA = 0, kOffB = kActBuf, kOffCond = 2 * kActBuf, kOffW = kOffCond + kCondRing * kCondStage;
constexpr uint32_t kWMain = 3 * 4096, kWAux = 8192, kW5 = 4096;           // image bytes: 3 taps x (KB 32 x NTp 32 x hi|lo), FiLM 1x1, c5
constexpr uint32_t kWBytes = 4 * kWMain + 2 * kWAux + kW5;
constexpr uint32_t kOffBar = kOffW + kWBytes;
constexpr uint32_t kBlockSmem = kOffBar + 256;
constexpr uint32_t kAccCols = 96;             // conv | FiLM scale | FiLM shift, 32 columns each
constexpr uint32_t kYCol0 = 2 * kAccCols;     // TMEM columns of the fp32 residual y: 32 per tile
constexpr int kBEpiWarps = 12, kBMmaWarp = 12, kBProdWarp = 13, kBThreads = 448;
static_assert(kBPad >= kBMaxDil && kBPad % 8 == 0 && kBSlots % 8 == 0, "slot geometry");
static_assert(kBPad + 7 + kBW + kBMaxDil <= kBSlots, "buffer too short");
static_assert(kOffB % 128 == 0 && kOffCond % 128 == 0 && kCondStage % 128 == 0 && kCondPlane % 128 == 0 && kActPlane % 128 == 0, "TMA alignment");
static_assert(kBlockSmem <= 227 * 1024, "shared memory");

struct alignas(64) UpBlockParams {
    CUtensorMap tm_p_hi, tm_p_lo, tm_c_hi, tm_c_lo;   // [3 chunks][row / 8][128 B] views; boxes of 73 / 17 row groups
    const bf16* w[5];
    uint32_t w_bytes[5], w_off16[5];                  // image sizes and their offsets (16-byte units) in the resident region
    const float* bias[5];
    const float* film_bias[2];
    const float* xi;
    float* xo;
    long long rows, n_seg;
    int T, segs_per_utt, xo_cs;
    int dil[4];
};

// Geometry of one segment (window) of the walk: seg = utterance * segs_per_utt + k.
struct SegGeo {
    long long baseT;   // operand row of the utterance's first time step
    long long R0;      // operand row held by buffer slot 0 (multiple of 8; negative / past the end reads as zeros)
    int w0;            // time step of window row 0 (k * 432 - 40; negative for the first window)
    int sh;            // buffer slot of window row 0
    __device__ SegGeo(const UpBlockParams& p, long long seg) {
        const long long bq = seg / p.segs_per_utt;
        const int k = (int)(seg - bq * p.segs_per_utt);
        baseT = bq * p.T;
        w0 = k * kBS - kBHalo;
        R0 = (baseT + w0 - kBPad) & ~7LL;
        sh = (int)(baseT + w0 - R0);
    }
};

__global__ void __launch_bounds__(kBThreads, 1) tc_up24_block_kernel(const __grid_constant__ UpBlockParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar = sb + kOffBar;
    const uint32_t wfull = bar, in_full = bar + 8, in_free = bar + 16, cond_full = bar + 24, cond_empty = bar + 40;
    const uint32_t acc_full = bar + 56, acc_empty = bar + 72, act_ready = bar + 88;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 128);

    if (tid == 0) {
        mbar_init(wfull, 1); mbar_init(in_full, 1); mbar_init(in_free, 1);
        for (uint32_t s = 0; s < kCondRing; ++s) { mbar_init(cond_full + 8 * s, 1); mbar_init(cond_empty + 8 * s, 1); }
        for (uint32_t b = 0; b < 2; ++b) { mbar_init(acc_full + 8 * b, 1); mbar_init(acc_empty + 8 * b, kBEpiWarps); }
        for (uint32_t j = 0; j < kBT; ++j) mbar_init(act_ready + 8 * j, kBEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // activation buffers and cond ring start as zeros: rows nobody writes and the aliased "fourth chunk" must be finite
    for (uint32_t i = tid; i < kOffW / 16; i += kBThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    if (warp == kBMmaWarp) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == kBProdWarp) {
        // ================= producer (one thread) =================
        if (lane == 0) {
            mbar_arrive_expect_tx(wfull, kWBytes);
            for (int l = 0; l < 5; ++l) bulk_g2s(sb + kOffW + (p.w_off16[l] << 4), p.w[l], p.w_bytes[l], wfull);
            uint32_t cs = 0, cph = 0;
            bool cwrapped = false;
            long long it = 0;
            for (long long seg = blockIdx.x; seg < p.n_seg; seg += gridDim.x, ++it) {
                const SegGeo g(p, seg);
                const uint32_t dst = sb + ((it & 1) ? kOffB : kOffA);
                if (it > 0) mbar_wait(in_free, (uint32_t)(it - 1) & 1u);       // c4 of the previous segment has read this buffer
                mbar_arrive_expect_tx(in_full, kActBuf);
                tma_load_3d(dst, &p.tm_p_hi, 0, (int)(g.R0 >> 3), 0, in_full);
                tma_load_3d(dst + kActPlane, &p.tm_p_lo, 0, (int)(g.R0 >> 3), 0, in_full);
                for (int pass = 0; pass < 2; ++pass) {                          // FiLM1 (c2), FiLM2 (c4)
                    for (int j = 0; j < kBT; ++j) {
                        const long long row0 = g.baseT + g.w0 + (long long)j * kBM;
                        const int grp = (int)((row0 - (row0 & 7)) >> 3);
                        const uint32_t full = cond_full + 8 * cs, cdst = sb + kOffCond + cs * kCondStage;
                        if (cwrapped) mbar_wait(cond_empty + 8 * cs, cph);
                        mbar_arrive_expect_tx(full, kCondStage);
                        tma_load_3d(cdst, &p.tm_c_hi, 0, grp, 0, full);
                        tma_load_3d(cdst + kCondPlane, &p.tm_c_lo, 0, grp, 0, full);
                        if (++cs == kCondRing) { cs = 0; cph = cwrapped ? cph ^ 1u : 0u; cwrapped = true; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kBMmaWarp) {
        // ================= MMA issue (one elected lane; loops are warp-uniform) =================
        const bool leader = elect_one() != 0;
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);          // SBO = 128 bytes, descriptor version 1
        const uint32_t idesc_m = umma_idesc(kBM, 32), idesc_x = umma_idesc(kBM, 64);
        const uint32_t lbo16 = kActLbo >> 4, plane16 = kActPlane >> 4, clbo16 = kCondLbo >> 4, cplane16 = kCondPlane >> 4;
        const uint32_t a16 = ((sb + kOffA) & 0x3FFFFu) >> 4, b16 = ((sb + kOffB) & 0x3FFFFu) >> 4;
        const uint32_t c16 = ((sb + kOffCond) & 0x3FFFFu) >> 4, w16 = ((sb + kOffW) & 0x3FFFFu) >> 4;
        mbar_wait(wfull, 0u);
        tc_fence_after();
        uint32_t tcount = 0, cs = 0, cph = 0;
        uint32_t seen[kBT];                                          // act_ready phases observed, per tile
#pragma unroll
        for (int j = 0; j < kBT; ++j) seen[j] = 0;
        long long it = 0;
        for (long long seg = blockIdx.x; seg < p.n_seg; seg += gridDim.x, ++it) {
            const SegGeo g(p, seg);
            const uint32_t in_off = (it & 1) ? kOffB : kOffA;
            const uint32_t in16 = (it & 1) ? b16 : a16, ot16 = (it & 1) ? a16 : b16;
            mbar_wait(in_full, (uint32_t)it & 1u);
            // replicate padding of the input for c1 (dil 1): t = -1 <- t = 0 and t = T <- t = T - 1, where the window has them
            if (lane < 12) {
                const int which = lane / 6, e = lane % 6;
                uint8_t* col = smem + in_off + (e / 3) * kActPlane + (e % 3) * kActLbo;
                if (which == 0 && g.w0 <= 0) {
                    const int s0 = g.sh - g.w0;
                    *reinterpret_cast<uint4*>(col + (s0 - 1) * 16) = *reinterpret_cast<const uint4*>(col + s0 * 16);
                }
                if (which == 1 && p.T - 1 < g.w0 + kBW) {
                    const int sT = g.sh + (p.T - 1 - g.w0);
                    *reinterpret_cast<uint4*>(col + (sT + 1) * 16) = *reinterpret_cast<const uint4*>(col + sT * 16);
                }
            }
            fence_proxy_async();
            __syncwarp();
            tc_fence_after();
#pragma unroll 1
            for (int l = 0; l < 5; ++l) {
                const uint32_t src16 = (l & 1) ? ot16 : in16;
                const uint32_t d = l < 4 ? (uint32_t)p.dil[l] : 0u;
                const bool film = l == 1 || l == 3;
                const uint32_t wl16 = w16 + p.w_off16[l];
#pragma unroll
                for (int j = 0; j < kBT; ++j) {
                    const uint32_t buf = tcount & 1u, buse = tcount >> 1;
                    if (buse > 0) mbar_wait(acc_empty + 8u * buf, (buse - 1) & 1u);      // epilogue drained this accumulator
                    if (l > 0) {
                        // the rows this tile reads (its own and up to 27 of each neighbour's) must have been published; every
                        // phase of a barrier is observed exactly once and in order (a parity wait cannot tell phase u from u - 2)
                        const uint32_t need = (uint32_t)it * 4u + (uint32_t)l;
#pragma unroll
                        for (int jj = 0; jj < kBT; ++jj) {
                            if (jj < j - 1 || jj > j + 1) continue;
                            while (seen[jj] < need) { mbar_wait(act_ready + 8u * jj, seen[jj] & 1u); ++seen[jj]; }
                        }
                    }
                    tc_fence_after();
                    const uint32_t dacc = tmem + buf * kAccCols;
                    const uint32_t a_lo0 = (lbo16 << 16) + src16 + (uint32_t)(g.sh + kBM * j) - d;   // tap 0 reads from row - dil
                    const uint32_t b_lo0 = (32u << 16) + wl16;
                    if (leader) {
                        if (l < 4) issue_stage<3, 2>(dacc, a_lo0, b_lo0, d, 256u, 2u * lbo16, 64u, plane16, 128u, desc_hi, idesc_m, 0u);
                        else issue_stage<1, 2>(dacc, a_lo0, b_lo0, 0u, 0u, 2u * lbo16, 64u, plane16, 128u, desc_hi, idesc_m, 0u);
                    }
                    if (film) {
                        mbar_wait(cond_full + 8u * cs, cph);
                        tc_fence_after();
                        const uint32_t off_x = (uint32_t)((g.baseT + g.w0 + (long long)j * kBM) & 7);
                        const uint32_t ax = (clbo16 << 16) + c16 + cs * (kCondStage >> 4) + off_x;
                        const uint32_t bx = (64u << 16) + wl16 + (kWMain >> 4);                       // FiLM image follows the 3 taps
                        if (leader) {
                            issue_stage<1, 2>(dacc + 32u, ax, bx, 0u, 0u, 2u * clbo16, 128u, cplane16, 256u, desc_hi, idesc_x, 0u);
                            umma_commit(cond_empty + 8u * cs);
                        }
                        if (++cs == kCondRing) { cs = 0; cph ^= 1u; }
                    }
                    if (leader) umma_commit(acc_full + 8u * buf);
                    ++tcount;
                }
                if (l == 3 && leader) umma_commit(in_free);          // `other` may take the next segment's input
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue: warp = (lane quarter, 8-channel group) =================
        const int quarter = warp & 3, cg = warp >> 2;
        const int rloc = quarter * 32 + lane;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        uint32_t tcount = 0;
        long long it = 0;
        for (long long seg = blockIdx.x; seg < p.n_seg; seg += gridDim.x, ++it) {
            const SegGeo g(p, seg);
            const uint32_t in_off = (it & 1) ? kOffB : kOffA, ot_off = (it & 1) ? kOffA : kOffB;
#pragma unroll 1
            for (int l = 0; l < 5; ++l) {
                const bool film = l == 1 || l == 3;
                const float* bias = p.bias[l] + cg * 8;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias)), b1 = __ldg(reinterpret_cast<const float4*>(bias) + 1);
                float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, h0 = s0, h1 = s0;
                if (film) {
                    const float* fb = p.film_bias[l >> 1] + cg * 8;                  // [scale bias (32) | shift bias (32)]
                    s0 = __ldg(reinterpret_cast<const float4*>(fb)); s1 = __ldg(reinterpret_cast<const float4*>(fb) + 1);
                    h0 = __ldg(reinterpret_cast<const float4*>(fb + 32)); h1 = __ldg(reinterpret_cast<const float4*>(fb + 32) + 1);
                }
                uint8_t* out = smem + ((l & 1) ? in_off : ot_off) + cg * kActLbo;   // L1, L3 -> other; L2, L4 -> in
                const int nd = l < 3 ? p.dil[l + 1] : 0;                             // replicate slots the next conv reads
                const int act = l == 3 ? TC_ACT_NONE : TC_ACT_LRELU;
                for (int j = 0; j < kBT; ++j, ++tcount) {
                    const uint32_t buf = tcount & 1u, buse = tcount >> 1;
                    const int r = j * kBM + rloc, t = g.w0 + r;
                    const bool inside = t >= 0 && t < p.T;
                    const long long row = g.baseT + t;
                    float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
                    if (l == 1 && inside) {                                          // residual x, requested before the wait
                        const float4* rp = reinterpret_cast<const float4*>(p.xi + ((long long)cg * p.rows + row) * 8);
                        r0 = __ldg(rp);
                        r1 = __ldg(rp + 1);
                    }
                    mbar_wait(acc_full + 8u * buf, buse & 1u);
                    tc_fence_after();
                    const uint32_t ta = tmem + buf * kAccCols + lane_sel + (uint32_t)(cg * 8);
                    const uint32_t ya = tmem + kYCol0 + (uint32_t)(j * 32 + cg * 8) + lane_sel;
                    float v[8], sc[8], sf[8], yv[8];
                    tmem_ld8(ta, v);
                    if (film) { tmem_ld8(ta + 32u, sc); tmem_ld8(ta + 64u, sf); }
                    if (l == 3) tmem_ld8(ya, yv);
                    tmem_ld_wait();
                    v[0] = __fadd_rn(v[0], b0.x); v[1] = __fadd_rn(v[1], b0.y); v[2] = __fadd_rn(v[2], b0.z); v[3] = __fadd_rn(v[3], b0.w);
                    v[4] = __fadd_rn(v[4], b1.x); v[5] = __fadd_rn(v[5], b1.y); v[6] = __fadd_rn(v[6], b1.z); v[7] = __fadd_rn(v[7], b1.w);
                    if (film) {
                        const float sbv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                        const float hbv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)                                   // FiLM: x * scale + shift (decoder.py:97)
                            v[i] = __fadd_rn(__fmul_rn(v[i], __fadd_rn(sc[i], sbv[i])), __fadd_rn(sf[i], hbv[i]));
                    }
                    if (l == 1) {
                        v[0] = __fadd_rn(v[0], r0.x); v[1] = __fadd_rn(v[1], r0.y); v[2] = __fadd_rn(v[2], r0.z); v[3] = __fadd_rn(v[3], r0.w);
                        v[4] = __fadd_rn(v[4], r1.x); v[5] = __fadd_rn(v[5], r1.y); v[6] = __fadd_rn(v[6], r1.z); v[7] = __fadd_rn(v[7], r1.w);
                        tmem_st8(ya, v);                                              // y stays in TMEM until c4's epilogue
                        tmem_st_wait();
                    }
                    if (l == 3) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(v[i], yv[i]);
                    }
                    if (l < 4) {
                        if (inside) {
                            uint32_t h[4], lw[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split2(apply_act(v[2 * i], act), apply_act(v[2 * i + 1], act), h[i], lw[i]);
                            const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            const int sl = g.sh + r;
                            *reinterpret_cast<uint4*>(out + sl * 16) = hv;
                            *reinterpret_cast<uint4*>(out + kActPlane + sl * 16) = lv;
                            if (t == 0) {                                            // replicate padding for the next conv's taps
                                for (int q = 1; q <= nd; ++q) {
                                    *reinterpret_cast<uint4*>(out + (sl - q) * 16) = hv;
                                    *reinterpret_cast<uint4*>(out + kActPlane + (sl - q) * 16) = lv;
                                }
                            }
                            if (t == p.T - 1) {
                                for (int q = 1; q <= nd; ++q) {
                                    *reinterpret_cast<uint4*>(out + (sl + q) * 16) = hv;
                                    *reinterpret_cast<uint4*>(out + kActPlane + (sl + q) * 16) = lv;
                                }
                            }
                        }
                        fence_proxy_async();                                         // the next layer's MMAs read these rows
                    } else if (inside && r >= kBHalo && r < kBHalo + kBS && cg * 8 + 8 <= p.xo_cs) {
                        float4* o = reinterpret_cast<float4*>(p.xo + ((long long)cg * p.rows + row) * 8);
                        o[0] = make_float4(v[0], v[1], v[2], v[3]);
                        o[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(acc_empty + 8u * buf);
                        if (l < 4) mbar_arrive(act_ready + 8u * (uint32_t)j);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBMmaWarp) tmem_dealloc(tmem, 512);
}

bool k3_ok(const TcConvW& c, bool film) {
    return c.taps == 3 && c.Cin == 24 && c.Cout == 24 && c.KB == 32 && c.nkb == 1 && c.NTp == 32 && c.n_tiles == 1 &&
           (film ? (c.aux_mode == TC_AUX_FILM && c.aux_cin == 24 && c.aux_nkb == 1 && c.tile_elems * 2 == kWMain + kWAux)
                 : (c.aux_mode == TC_AUX_NONE && c.tile_elems * 2 == kWMain));
}

}  // namespace

bool tc_up24_block_supported(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5) {
    return k3_ok(c1, false) && k3_ok(c2, true) && k3_ok(c3, false) && k3_ok(c4, true) && c5.taps == 1 && c5.Cin == 24 &&
           c5.Cout <= 24 && c5.KB == 32 && c5.nkb == 1 && c5.NTp == 32 && c5.n_tiles == 1 && c5.aux_mode == TC_AUX_NONE &&
           c5.tile_elems * 2 == kW5;
}

int tc_up24_block_launch(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5,
                         const TcUpBlockArgs& a, cudaStream_t s) {
    TVC_REQUIRE(tc_up24_block_supported(c1, c2, c3, c4, c5), "tc_up24_block: conv shapes are not the 24-channel Upsample block");
    TVC_REQUIRE(a.p_hi && a.p_lo && a.c_hi && a.c_lo && a.xi && a.xo && a.B > 0 && a.T > 0, "tc_up24_block: missing argument");
    TVC_REQUIRE(a.xo_cs % 8 == 0 && a.xo_cs >= 8, "tc_up24_block: output capacity %d", a.xo_cs);
    for (int i = 0; i < 4; ++i)
        TVC_REQUIRE(a.dil[i] >= 1 && a.dil[i] <= kBMaxDil, "tc_up24_block: dilation %d out of range", a.dil[i]);
    TVC_REQUIRE(a.dil[0] == 1 && a.dil[0] + a.dil[1] + a.dil[2] + a.dil[3] <= kBHalo, "tc_up24_block: dilations exceed the window halo");
    static PerDeviceOnce attr;
    TVC_TRY(attr.run([] { TVC_CUDA(cudaFuncSetAttribute(tc_up24_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlockSmem)); return 0; }));
    UpBlockParams p;
    memset(&p, 0, sizeof(p));
    p.rows = (long long)a.B * a.T;
    TVC_TRY(tc_make_plane_map(&p.tm_p_hi, a.p_hi, p.rows, 3, 3, kBSlots / 8));
    TVC_TRY(tc_make_plane_map(&p.tm_p_lo, a.p_lo, p.rows, 3, 3, kBSlots / 8));
    TVC_TRY(tc_make_plane_map(&p.tm_c_hi, a.c_hi, p.rows, 3, 3, kCondG));
    TVC_TRY(tc_make_plane_map(&p.tm_c_lo, a.c_lo, p.rows, 3, 3, kCondG));
    const TcConvW* cw[5] = {&c1, &c2, &c3, &c4, &c5};
    uint32_t off = 0;
    for (int l = 0; l < 5; ++l) {
        p.w[l] = cw[l]->w;
        p.w_bytes[l] = (uint32_t)(cw[l]->tile_elems * sizeof(bf16));
        p.w_off16[l] = off >> 4;
        off += p.w_bytes[l];
        p.bias[l] = cw[l]->bias;
    }
    TVC_REQUIRE(off == kWBytes, "tc_up24_block: weight images total %u bytes, expected %u", off, kWBytes);
    p.film_bias[0] = c2.film_bias;
    p.film_bias[1] = c4.film_bias;
    p.xi = a.xi; p.xo = a.xo; p.xo_cs = a.xo_cs;
    p.T = a.T;
    p.segs_per_utt = cdiv(a.T, kBS);
    p.n_seg = (long long)a.B * p.segs_per_utt;
    for (int i = 0; i < 4; ++i) p.dil[i] = a.dil[i];
    int dev = 0, sms = 148;
    TVC_CUDA(cudaGetDevice(&dev));
    TVC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned grid = (unsigned)(p.n_seg < sms ? p.n_seg : sms);
    tc_up24_block_kernel<<<grid, kBThreads, kBlockSmem, s>>>(p);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc

// Fused Upsample block at the FilterNet's highest rate (24 channels) + the output layer, sm_100a.
//
//   x  = interp(x4, x5)                 F.interpolate(scale_factor=5, mode='linear')   (module/tinyvc/decoder.py:174)
//   p  = lrelu(x)
//   h1 = lrelu(c1(p))                   k = 3, dil 1
//   y  = FiLM1(c2(h1); cond) + x        k = 3, dil 3          (decoder.py:165-171, FiLM :88-97)
//   h3 = lrelu(c3(lrelu(y)))            k = 3, dil 9
//   z  = FiLM2(c4(h3); cond) + y        k = 3, dil 27
//   xo = c5(z)                          1 x 1
//   out = output_layer(xo)              k = 7, replicate pad 3, 24 -> 1                (decoder.py:220,233)
//
// Unfused, this chain moves ~1.9 KB per time row through L2 / HBM (the x5 resampler writes the fp32 residual and the
// activation planes, every 24-channel intermediate is written as split planes and read back, the output conv re-reads xo)
// and is bandwidth-bound.  Here a CTA owns a WINDOW of 512 rows of one utterance (4 MMA tiles of 128 rows), builds the
// resampled input in shared memory from the low-rate tensor (19 B per row instead of 192), keeps every intermediate in
// shared memory as ready-made UMMA operands (chunk-major split planes are exactly the K-major smem order), the fp32
// residual `y` in TMEM and xo in shared memory, and produces the 426 waveform samples in the middle of the window: the
// 43 rows on either side (1 + 3 + 9 + 27 for the convs, 3 for the output layer) are recomputed by the neighbouring
// windows.  Per row it reads x4 (1/5 row), cond (twice, the second time from L2) and writes 4 bytes.
//
// The arithmetic (resampler formula, MMA order per tile, epilogue operations, the output conv's summation order) is that
// of interp_cl / tc_conv.cu / out_conv_k7_cl, so the result is bit-identical to the separate launches
// (tests/test_gpu_fused_block.py compares the two).
//
// Shared memory (bytes):   [buffer A: hi|lo planes, 3 chunks x 584 slots x 16 B][cond ring: 2 x (hi|lo, 3 chunks x 136 slots)]
//   [buffer B][weight images of c1..c5, resident][output-layer weights][mbarriers].  Window row r lives in slot 32 + r.
//   Activations have 24 channels but a K-step is 16: the second K-step's upper chunk aliases the first chunk of
//   whatever follows (lo plane, cond ring, weights -- all finite bf16, zero-initialised) and meets zero weights.
// Warp roles (576 threads): warps 0-11 epilogue (4 TMEM lane quarters x 3 column groups) and the output conv, warp 12 MMA
//   issue, warp 13 TMA producer (weights once, a cond tile per FiLM tile), warps 14-17 build the next window's input
//   (resample, leaky-ReLU, split) in the idle buffer as soon as c4's MMAs have retired (in_free), i.e. behind c5, its
//   epilogues and the output conv.
// Layer l + 1 of tile j starts once the epilogues of layer l for tiles j - 1, j, j + 1 have published their rows
// (act_ready[j] mbarriers); buffers ping-pong (L1: in -> other, L2: other -> in, ...).  c5's epilogue stores xo (fp32) over
// the rows of `in` its own MMA has finished reading; the output conv runs between two named barriers of the epilogue warps.
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
constexpr int kBConvHalo = 40;                // 1 + 3 + 9 + 27: rows at either end of a window where c1..c4 lack their taps
constexpr int kBHalo = kBConvHalo + 3;        // + the output layer's 3: rows a window cannot produce on either side
constexpr int kBS = kBW - 2 * kBHalo;         // samples a window produces (426)
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
// Order matters: the K-step that covers channels 16-31 reads a fourth chunk that does not exist (24 channels), i.e. the
// first chunk of whatever follows the plane at the same row slots -- for a hi plane its own lo plane, for buffer A's lo plane
// the cond ring, for buffer B's lo plane the weight images: always finite bf16 (it meets zero weights, but 0 x NaN = NaN).
// The same holds for a cond stage (its lo plane is followed by the next stage, the last one by kCondTail zeros).
// Neither a buffer nor the cond ring may be followed directly by a buffer: at the end of a segment a buffer holds fp32 xo.
constexpr uint32_t kCondTail = 18 * 128;      // >= kCondLbo, zero for the life of the kernel
constexpr uint32_t kOffA = 0, kOffCond = kActBuf, kOffB = kOffCond + kCondRing * kCondStage + kCondTail, kOffW = kOffB + kActBuf;
static_assert(kCondTail >= kCondLbo, "the alias of the last cond stage's fourth chunk must stay inside the zero tail");
constexpr uint32_t kWMain = 3 * 4096, kWAux = 8192, kW5 = 4096;           // image bytes: 3 taps x (KB 32 x NTp 32 x hi|lo), FiLM 1x1, c5
constexpr uint32_t kWBytes = 4 * kWMain + 2 * kWAux + kW5;
constexpr uint32_t kOffOutW = kOffW + kWBytes;                            // output layer: [7 taps][24 channels] fp32, then the bias
constexpr uint32_t kOffBar = kOffOutW + 1024;
constexpr uint32_t kBlockSmem = kOffBar + 256;
constexpr uint32_t kAccCols = 128;            // conv x_hi*w_hi + x_lo*w_hi | conv x_hi*w_lo | FiLM scale | FiLM shift, 32 columns each
#ifndef TVC_BLK_BUFS
#define TVC_BLK_BUFS 3
#endif
#ifndef TVC_BLK_EARLY
#define TVC_BLK_EARLY 1
#endif
constexpr uint32_t kAccBufs = TVC_BLK_BUFS;   // accumulator ring: tile-op k uses buffer k % kAccBufs (3 x 128 + 4 x 32 y columns = the 512 of TMEM)
constexpr uint32_t kYCol0 = kAccBufs * kAccCols;   // TMEM columns of the fp32 residual y: 32 per tile
static_assert(kYCol0 + 4 * 32 <= 512, "TMEM columns");
constexpr int kBEpiWarps = 12, kBMmaWarp = 12, kBProdWarp = 13, kBLoadWarp0 = 14, kBLoadWarps = 4, kBThreads = 576;
constexpr int kBEpiThreads = kBEpiWarps * 32;
constexpr int kBInRows = kBW + 2;             // rows -1 .. 512 of the window: what c1 (dil 1) reads
static_assert(kBPad >= kBMaxDil && kBPad % 8 == 0 && kBSlots % 8 == 0, "slot geometry");
static_assert(kBPad + 7 + kBW + kBMaxDil <= kBSlots, "buffer too short");
static_assert(kOffB % 128 == 0 && kOffCond % 128 == 0 && kCondStage % 128 == 0 && kCondPlane % 128 == 0 && kActPlane % 128 == 0, "TMA alignment");
static_assert(kBlockSmem <= 227 * 1024, "shared memory");

struct alignas(64) UpBlockParams {
    CUtensorMap tm_c_hi, tm_c_lo;                     // [3 chunks][row / 8][128 B] views of the cond planes; boxes of 17 row groups
    const bf16* w[5];
    uint32_t w_bytes[5], w_off16[5];                  // image sizes and their offsets (16-byte units) in the resident region
    const float* bias[5];
    const float* film_bias[2];
    const float* x4;                                  // block input before resampling: fp32 chunk-major, B * T4 rows x 24 channels
    const float *out_w, *out_b;                       // output_layer.weight [1][24][7], .bias [1]
    float* out;                                       // waveform [B][T]
    long long rows, rows4, n_seg;
    int T, T4, segs_per_utt, seg_lo;                   // windows walked per utterance and the index of the first one (output pruning)
    int T4c, x4_off;                                   // x4 holds rows [x4_off, x4_off + T4c) of every utterance's low-rate tensor
    float scale;                                      // resampler scale (float)(1 / 5)
    int dil[4];
};

// Geometry of one segment (window) of the walk: seg = utterance * segs_per_utt + k.
struct SegGeo {
    long long baseT;   // output row of the utterance's first time step
    long long base4;   // row of its first time step in the low-rate input
    int w0;            // time step of window row 0 (k * 426 - 43; negative for the first window)
    static constexpr int sh = kBPad;   // buffer slot of window row 0
    __device__ SegGeo(const UpBlockParams& p, long long seg) {
        const long long bq = seg / p.segs_per_utt;
        const int k = p.seg_lo + (int)(seg - bq * p.segs_per_utt);
        baseT = bq * p.T;
        base4 = bq * p.T4c - p.x4_off;                   // so that base4 + (row of the full-length tensor) addresses the window
        w0 = k * kBS - kBHalo;
    }
};

// 8 channels (chunk q) of the resampled block input at time t of the utterance whose low-rate rows start at base4:
// interp_cl's arithmetic (tc_frame.cu), so the values are the ones the separate resampler launch produces.
__device__ __forceinline__ void resample8(const UpBlockParams& p, long long base4, int q, int t, float v[8]) {
    const LinCoord c = lin_coord(t, p.scale, p.T4);
    const float4* p0 = reinterpret_cast<const float4*>(p.x4 + ((long long)q * p.rows4 + base4 + c.i0) * 8);
    const float4* p1 = reinterpret_cast<const float4*>(p.x4 + ((long long)q * p.rows4 + base4 + c.i1) * 8);
    const float4 x0 = __ldg(p0), x1 = __ldg(p0 + 1), z0 = __ldg(p1), z1 = __ldg(p1 + 1);
    v[0] = lin_blend(x0.x, z0.x, c); v[1] = lin_blend(x0.y, z0.y, c); v[2] = lin_blend(x0.z, z0.z, c); v[3] = lin_blend(x0.w, z0.w, c);
    v[4] = lin_blend(x1.x, z1.x, c); v[5] = lin_blend(x1.y, z1.y, c); v[6] = lin_blend(x1.z, z1.z, c); v[7] = lin_blend(x1.w, z1.w, c);
}

__global__ void __launch_bounds__(kBThreads, 1) tc_up24_block_kernel(const __grid_constant__ UpBlockParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar = sb + kOffBar;
    const uint32_t wfull = bar, in_full = bar + 8, in_free = bar + 16, cond_full = bar + 24, cond_empty = bar + 40;
    const uint32_t acc_full = bar + 56, acc_empty = acc_full + 8 * kAccBufs, act_ready = acc_empty + 8 * kAccBufs;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffBar + 192);
    static_assert(56 + 16 * kAccBufs + 8 * kBT <= 192, "barrier block");

    if (tid == 0) {
        mbar_init(wfull, 1); mbar_init(in_full, kBLoadWarps); mbar_init(in_free, 1);
        for (uint32_t s = 0; s < kCondRing; ++s) { mbar_init(cond_full + 8 * s, 1); mbar_init(cond_empty + 8 * s, 1); }
        for (uint32_t b = 0; b < kAccBufs; ++b) { mbar_init(acc_full + 8 * b, 1); mbar_init(acc_empty + 8 * b, kBEpiWarps); }
        for (uint32_t j = 0; j < kBT; ++j) mbar_init(act_ready + 8 * j, kBEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // activation buffers and cond ring start as zeros: rows nobody writes and the aliased "fourth chunk" must be finite
    for (uint32_t i = tid; i < kOffW / 16; i += kBThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    // output layer weights, re-laid [tap][channel] (torch keeps [channel][tap]), bias behind them
    for (int e = tid; e < 24 * 7 + 1; e += kBThreads) {
        float* ws = reinterpret_cast<float*>(smem + kOffOutW);
        if (e < 24 * 7) ws[(e % 7) * 24 + e / 7] = __ldg(p.out_w + e);
        else ws[e] = __ldg(p.out_b);
    }
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
        const uint32_t idesc_m = umma_idesc(kBM, 32), idesc_x = umma_idesc(kBM, 64);      // idesc_x: N = 64 (FiLM scale | shift; cat main)
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
            const uint32_t in16 = (it & 1) ? b16 : a16, ot16 = (it & 1) ? a16 : b16;
            mbar_wait(in_full, (uint32_t)it & 1u);             // the loader warps have built this window's input (replicate rows included)
            tc_fence_after();
#pragma unroll 1
            for (int l = 0; l < 5; ++l) {
                const uint32_t src16 = (l & 1) ? ot16 : in16;
                const uint32_t d = l < 4 ? (uint32_t)p.dil[l] : 0u;
                const bool film = l == 1 || l == 3;
                const uint32_t wl16 = w16 + p.w_off16[l];
#pragma unroll
                for (int j = 0; j < kBT; ++j) {
                    const uint32_t buf = tcount % kAccBufs, buse = tcount / kAccBufs;
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
                    // "cat" weight images (tc_conv.cuh): chunk stride 64 rows (hi 32 | lo 32), two MMAs per K-step
                    const uint32_t b_lo0 = (64u << 16) + wl16;
                    if (leader) {
                        if (l < 4) issue_stage_cat<3, 2>(dacc, a_lo0, b_lo0, d, 256u, 2u * lbo16, 128u, plane16, desc_hi, idesc_m, idesc_x, 0u);
                        else issue_stage_cat<1, 2>(dacc, a_lo0, b_lo0, 0u, 0u, 2u * lbo16, 128u, plane16, desc_hi, idesc_m, idesc_x, 0u);
                    }
                    if (film) {
                        mbar_wait(cond_full + 8u * cs, cph);
                        tc_fence_after();
                        const uint32_t off_x = (uint32_t)((g.baseT + g.w0 + (long long)j * kBM) & 7);
                        const uint32_t ax = (clbo16 << 16) + c16 + cs * (kCondStage >> 4) + off_x;
                        const uint32_t bx = (64u << 16) + wl16 + (kWMain >> 4);                       // FiLM image follows the 3 taps
                        if (leader) {
                            issue_stage<1, 2>(dacc + 64u, ax, bx, 0u, 0u, 2u * clbo16, 128u, cplane16, 256u, desc_hi, idesc_x, 0u);
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
    } else if (warp >= kBLoadWarp0) {
        // ================= input builders: p = lrelu(interp(x4)) as split planes, window rows -1 .. 512 =================
        // One (row, chunk) item = 8 channels: two low-rate rows in (L1 / L2 hits: 5 neighbours share them), 32 B of planes
        // out; consecutive lanes take consecutive rows (conflict-free 16-byte shared stores).  Time steps are clamped to the
        // utterance, which is c1's replicate padding.
        const int lt = tid - kBLoadWarp0 * 32;
        constexpr int kItems = 3 * kBInRows, kLT = kBLoadWarps * 32;
        long long it = 0;
        for (long long seg = blockIdx.x; seg < p.n_seg; seg += gridDim.x, ++it) {
            const SegGeo g(p, seg);
            uint8_t* dst = smem + ((it & 1) ? kOffB : kOffA);
            if (it > 0) mbar_wait(in_free, (uint32_t)(it - 1) & 1u);           // c4 of the previous segment has read this buffer
#pragma unroll 4
            for (int idx = lt; idx < kItems; idx += kLT) {
                const int q = idx / kBInRows, r = idx - q * kBInRows - 1;
                int t = g.w0 + r;
                t = t < 0 ? 0 : (t > p.T - 1 ? p.T - 1 : t);
                float v[8];
                resample8(p, g.base4, q, t, v);
                uint32_t h[4], l[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split2(leaky01(v[2 * i]), leaky01(v[2 * i + 1]), h[i], l[i]);
                uint8_t* o = dst + q * kActLbo + (SegGeo::sh + r) * 16;
                *reinterpret_cast<uint4*>(o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(o + kActPlane) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();                                               // c1's MMAs read these rows
            __syncwarp();
            if (lane == 0) mbar_arrive(in_full);
        }
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
                    const uint32_t buf = tcount % kAccBufs, buse = tcount / kAccBufs;
                    const int r = j * kBM + rloc, t = g.w0 + r;
                    const bool inside = t >= 0 && t < p.T;
                    const long long row = g.baseT + t;
                    float xr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (l == 1 && inside) resample8(p, g.base4, cg, t, xr);          // residual x = interp(x4), loads issued before the wait
                    mbar_wait(acc_full + 8u * buf, buse & 1u);
                    tc_fence_after();
                    const uint32_t ta = tmem + buf * kAccCols + lane_sel + (uint32_t)(cg * 8);
                    const uint32_t ya = tmem + kYCol0 + (uint32_t)(j * 32 + cg * 8) + lane_sel;
                    float v[8], v2[8], sc[8], sf[8], yv[8];
                    tmem_ld8(ta, v);
                    tmem_ld8(ta + 32u, v2);                                          // x_hi * w_lo columns
                    if (film) { tmem_ld8(ta + 64u, sc); tmem_ld8(ta + 96u, sf); }
                    if (l == 3) tmem_ld8(ya, yv);
                    tmem_ld_wait();
#if TVC_BLK_EARLY
                    // the accumulator is in registers: hand the buffer back before the arithmetic and the shared-memory stores
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + 8u * buf);
#endif
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(v[i], v2[i]);
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
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = __fadd_rn(v[i], xr[i]);
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
                            const int sl = SegGeo::sh + r;
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
                    } else if (inside) {
                        // xo (fp32) over the rows of `in` that this tile's own MMA has finished reading (c5 is 1 x 1): channels
                        // 0-3 of the chunk where the hi plane's row was, 4-7 where the lo plane's was
                        uint8_t* o = smem + in_off + cg * kActLbo + (SegGeo::sh + r) * 16;
                        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                        *reinterpret_cast<float4*>(o + kActPlane) = make_float4(v[4], v[5], v[6], v[7]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
#if !TVC_BLK_EARLY
                        mbar_arrive(acc_empty + 8u * buf);
#endif
                        if (l < 4) mbar_arrive(act_ready + 8u * (uint32_t)j);
                    }
                }
            }
            // ---- output layer (decoder.py:220,233): k = 7, replicate pad 3, 24 -> 1, out_conv_k7_cl's summation order.
            //      xo of all four tiles must be in place; the next segment's first epilogue overwrites this buffer.
            asm volatile("bar.sync 1, %0;" ::"n"(kBEpiThreads) : "memory");
            {
                const float* ws = reinterpret_cast<const float*>(smem + kOffOutW);
                const uint8_t* xs = smem + in_off;
                for (int r = kBHalo + tid; r < kBHalo + kBS; r += kBEpiThreads) {
                    const int t = g.w0 + r;
                    if (t < 0 || t >= p.T) continue;
                    float acc = ws[24 * 7];
#pragma unroll
                    for (int jt = 0; jt < 7; ++jt) {
                        int tt = t + jt - 3;
                        tt = tt < 0 ? 0 : (tt > p.T - 1 ? p.T - 1 : tt);
                        const uint8_t* row = xs + (SegGeo::sh + (tt - g.w0)) * 16;
#pragma unroll
                        for (int q = 0; q < 6; ++q) {
                            const float4 xv = *reinterpret_cast<const float4*>(row + (q >> 1) * kActLbo + (q & 1) * kActPlane);
                            const float4 wv = *reinterpret_cast<const float4*>(ws + jt * 24 + q * 4);
                            acc = fmaf(wv.x, xv.x, acc);
                            acc = fmaf(wv.y, xv.y, acc);
                            acc = fmaf(wv.z, xv.z, acc);
                            acc = fmaf(wv.w, xv.w, acc);
                        }
                    }
                    p.out[g.baseT + t] = acc;
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kBEpiThreads) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBMmaWarp) tmem_dealloc(tmem, 512);
}

bool k3_ok(const TcConvW& c, bool film) {
    return c.cat && c.taps == 3 && c.Cin == 24 && c.Cout == 24 && c.KB == 32 && c.nkb == 1 && c.NTp == 32 && c.n_tiles == 1 &&
           (film ? (c.aux_mode == TC_AUX_FILM && c.aux_cin == 24 && c.aux_nkb == 1 && c.tile_elems * 2 == kWMain + kWAux)
                 : (c.aux_mode == TC_AUX_NONE && c.tile_elems * 2 == kWMain));
}

}  // namespace

bool tc_up24_block_supported(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5) {
    return k3_ok(c1, false) && k3_ok(c2, true) && k3_ok(c3, false) && k3_ok(c4, true) && c5.cat && c5.taps == 1 && c5.Cin == 24 &&
           c5.Cout <= 24 && c5.KB == 32 && c5.nkb == 1 && c5.NTp == 32 && c5.n_tiles == 1 && c5.aux_mode == TC_AUX_NONE &&
           c5.tile_elems * 2 == kW5;
}

int tc_up24_block_launch(const TcConvW& c1, const TcConvW& c2, const TcConvW& c3, const TcConvW& c4, const TcConvW& c5,
                         const TcUpBlockArgs& a, cudaStream_t s) {
    TVC_REQUIRE(tc_up24_block_supported(c1, c2, c3, c4, c5), "tc_up24_block: conv shapes are not the 24-channel Upsample block");
    TVC_REQUIRE(a.x4 && a.c_hi && a.c_lo && a.out_w && a.out_b && a.out && a.B > 0 && a.T > 0, "tc_up24_block: missing argument");
    TVC_REQUIRE(a.T4 > 0 && a.T == 5 * a.T4, "tc_up24_block: the block resamples x5 (T = %d, T4 = %d)", a.T, a.T4);
    for (int i = 0; i < 4; ++i)
        TVC_REQUIRE(a.dil[i] >= 1 && a.dil[i] <= kBMaxDil, "tc_up24_block: dilation %d out of range", a.dil[i]);
    TVC_REQUIRE(a.dil[0] == 1 && a.dil[0] + a.dil[1] + a.dil[2] + a.dil[3] <= kBConvHalo, "tc_up24_block: dilations exceed the window halo");
    static PerDeviceOnce attr;
    TVC_TRY(attr.run([] { TVC_CUDA(cudaFuncSetAttribute(tc_up24_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBlockSmem)); return 0; }));
    UpBlockParams p;
    memset(&p, 0, sizeof(p));
    p.rows = (long long)a.B * a.T;
    const int T4c = a.x4_rows > 0 ? a.x4_rows : a.T4;
    TVC_REQUIRE(a.x4_off >= 0 && a.x4_off + T4c <= a.T4, "tc_up24_block: input window [%d, %d) outside the %d low-rate rows", a.x4_off, a.x4_off + T4c, a.T4);
    p.rows4 = (long long)a.B * T4c;
    p.T4c = T4c; p.x4_off = a.x4_off;
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
    p.x4 = a.x4; p.out_w = a.out_w; p.out_b = a.out_b; p.out = a.out;
    p.T = a.T; p.T4 = a.T4; p.scale = a.scale;
    const int t_hi = a.t_hi < 0 ? a.T : a.t_hi;
    TVC_REQUIRE(a.t_lo >= 0 && a.t_lo < t_hi && t_hi <= a.T, "tc_up24_block: output range [%d, %d) outside [0, %d)", a.t_lo, a.t_hi, a.T);
    p.seg_lo = a.t_lo / kBS;                                  // window k produces samples [k * 426, (k + 1) * 426)
    p.segs_per_utt = (t_hi - 1) / kBS - p.seg_lo + 1;
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

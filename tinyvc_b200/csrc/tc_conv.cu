// Dense Conv1d on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), sm_100a.
//
// One CTA computes a [128 time rows] x [NT output channels] tile of
//     y[row, co] = epi( bias[co] + sum_{tap,ci} W[co,ci,tap] * x[clamp_T(row + (tap-(taps-1)/2)*dil), ci]  (+ aux) )
// for the decoder's dense convolutions (module/tinyvc/decoder.py:93-94 FiLM 1x1s, :119-124 SourceNet
// heads, :143-146 Downsample, :165-171 Upsample, :201 content_in, :206 downs.0) as an implicit GEMM:
//     M = 128 rows (time), N = NT channels, K = taps * Cin, processed in K-stages of KB channels of one tap.
//
// Operands are chunk-major split planes (tc_conv.cuh): every 8-channel chunk is a dense [rows][16 B] array, which is also
// the order of the UMMA K-major smem operand ([chunk][row][16 B]).
//
// Warp roles (576 threads, one persistent CTA per SM walking tiles):
//   warps 13-16  producers.  A K-stage's activation window comes by ONE TMA tensor copy per plane
//                (cp.async.bulk.tensor.3d over the plane viewed as [chunk][row / 8][128 B]); k = 3 convs share one
//                (128 + 2*dil)-row window across the taps.  Windows that leave their utterance (replicate padding) and
//                k = 3 stages of tiles that straddle utterances are gathered per thread with 16-byte cp.async instead.
//                The stage's weight image (pre-packed in smem order) comes by cp.async.bulk.  Everything completes on the
//                stage's `full` mbarrier (complete_tx / cp.async.mbarrier.arrive.noinc).
//   warps 12,17  MMA issuers (tiles alternate; each owns half of the smem ring and one TMEM accumulator): lane 0 waits
//                `full`, issues 3 x KB/16 tcgen05.mma per tap (hi*hi, hi*lo, lo*hi; fp32 accumulate in TMEM),
//                tcgen05.commit -> the stage's `empty` mbarrier; after the last stage commit -> `acc_full`.
//   warps 0-11   epilogue: tcgen05.ld from the double-buffered TMEM accumulators, bias, FiLM (x*scale+shift from a
//                second accumulator fed by the aux 1x1 on the skip tensor), residual, activation, fp32 and/or re-split
//                bf16 planes, chunk-major stores.
// Every mbarrier wait is bounded (trap after ~2 s) so a protocol bug cannot hang the GPU.
#include <cuda.h>      // CUtensorMap (types only; cuTensorMapEncodeTiled is resolved through the runtime, no -lcuda)
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_conv.cuh"
#include "tc_ptx.cuh"

namespace tvc {

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 576;   // 12 epilogue warps + MMA warp + 4 producer warps + second MMA warp

struct alignas(64) TcKParams {
    // TMA tensor maps of the operand planes, each viewed as [chunk][row / 8][8 rows x 8 channels = 128 B]; the box is
    // [KB/8 chunks][G row groups][128 B] and lands in shared memory as [chunk][row][16 B], the UMMA K-major order
    CUtensorMap tm_a_hi, tm_a_lo, tm_x_hi, tm_x_lo;
    uint32_t w_bytes;              // > 0: the whole weight image of the (single) channel tile stays resident in shared memory
                                   //      (loaded once per CTA) instead of travelling with every K-stage
    int mma2;                      // 1: two MMA-issuing warps, each with its own half of the smem ring (tiles alternate)
    int tma_main, tma_aux;         // which stage types may use them (interior tiles only: no replicate padding inside)
    uint32_t lbo_main, lbo_aux;    // bytes between consecutive 8-channel chunk columns of a main / aux stage (= 128 * G)
    const bf16 *a_hi, *a_lo, *x_hi, *x_lo, *w;
    const float *bias, *film_bias, *res;
    float* y32;
    bf16 *y_hi, *y_lo;
    long long rows;                // B * T: rows of the aux input, the residual and the outputs
    long long rows_a;              // rows of the main input planes (= rows, or B * (T + 2 * a_pad) in padded mode)
    long long rows_y;              // rows of the plane output (B * (T + 2 * y_pad))
    int a_pad, y_pad, Tp;          // padded mode: replicate rows stored on either side of every utterance; Tp = T + 2 * a_pad
    long long tile_elems;
    int a_cs, x_cs, res_cs, y32_cs, y_cs;
    int T, dil, taps, nkb, aux_nkb, aux_mode, KB, NT, NTp, Cout;
    int ring;                      // smem ring depth in use
    int ring_alloc;                // slots laid out in shared memory (>= ring)
    long long row_tiles;
    int n_tiles;
    int halo;                      // 1: one stage holds a (128 + 2*dil)-row window shared by the 3 taps (tiles never straddle utterances)
                                   // 2: the same over an input whose replicate padding is STORED (a_pad >= dil rows on either side of
                                   //    every utterance): tiles walk the padded row space, every window is one contiguous tensor copy
    int R;                         // rows per A stage (128, or 128 + 2*dil in halo mode)
    int tiles_per_utt;             // halo mode
    uint32_t a_stage_bytes, b_stage_bytes;
    uint32_t tmem_cols;
    int epi_act, out_act;
    // fused top-k (kNN screening, SPEC 8): no output tensor; every row keeps its topk_c best columns while its CTA sweeps the
    // channel tiles (tile order is row-tile major), then writes their indices and a "margin too thin" flag
    int* topk_cand;                // [rows][kTopkC] column indices, best first (-1: fewer than kTopkC valid columns)
    int* topk_flag;                // [rows] 1 when score[k-1] - score[kTopkC-1] <= topk_eps (a true member may have been missed)
    int topk_k, topk_n;            // k of the final selection; number of real columns (channels >= topk_n are padding)
    float topk_eps;
    float* topk_score;             // split sweep: [rows][topk_splits][kTopkC] scores next to the indices
    int topk_splits, topk_nper;    // ranges the channel tiles are cut into, channel tiles per range
    uint32_t topk_smem;            // byte offset of the merge scratch inside dynamic shared memory
    int row_major;                 // 1: tile id = row_tile * n_tiles + n_tile, whole row tiles per CTA
    // fused down-resampler (SPEC 9): besides its own outputs the epilogue writes F.interpolate(y, scale_factor=1/dec_f,
    // mode='linear') of the conv result as the next Downsample block's operands (decoder.py:148): raw planes (down_res input)
    // and leaky-ReLU'd planes (c1 input, optionally with dec_pad stored replicate rows)
    int dec_f, dec_pad;
    float dec_scale;               // interp_cl's source-index scale, (float)dec_f
    bf16 *dec_r_hi, *dec_r_lo, *dec_a_hi, *dec_a_lo;
    int dbg;   // ablation switches for profiling (TVC_TC_DBG): 1 no loads, 2 no MMAs, 4 no epilogue math/stores
    uint2* trace;   // developer timeline (TVC_TC_TRACE): CTA 0 logs {clock, role|event|tile|stage} per pipeline event
};

// ---------------------------------------------------------------------------------------------
// Epilogue specialisations.  The element-wise tail is the instruction-bound part of the small-channel
// convs, so the combinations the decoder uses are compiled with their flags as constants; the generic
// instantiation (flags read from the parameters) serves everything else (parity probes).
// ---------------------------------------------------------------------------------------------
struct EpiSpec { int film, res, y32, planes, epi_act, out_act, topk, dec; };
constexpr int kNumSpecs = 10;
constexpr int kTopkC = 8;          // candidates kept per row by the fused top-k epilogue
__host__ __device__ constexpr EpiSpec epi_spec(int i) {
    return i == 9 ? EpiSpec{0, 0, 0, 1, TC_ACT_NONE, TC_ACT_NONE, 0, 1}   // skip tensor + the next Downsample block's resampled operands
         : i == 8 ? EpiSpec{0, 0, 0, 0, TC_ACT_NONE, TC_ACT_NONE, 1}   // kNN screening: per-row top candidates, no output tensor
         : i == 0 ? EpiSpec{0, 0, 1, 0, TC_ACT_NONE, TC_ACT_NONE}      // plain fp32 output
         : i == 1 ? EpiSpec{0, 0, 1, 0, TC_ACT_GELU, TC_ACT_NONE}      // ConvNeXt c2
         : i == 2 ? EpiSpec{0, 1, 1, 1, TC_ACT_NONE, TC_ACT_NONE}      // ConvNeXt c3 (+residual)
         : i == 3 ? EpiSpec{0, 0, 1, 0, TC_ACT_ELU1, TC_ACT_NONE}      // SourceNet heads
         : i == 4 ? EpiSpec{0, 0, 1, 1, TC_ACT_NONE, TC_ACT_NONE}      // skip tensors (downs.0, Downsample c3)
         : i == 5 ? EpiSpec{0, 0, 0, 1, TC_ACT_NONE, TC_ACT_LRELU}     // c1/c3 of Up, c1/c2 of Down
         : i == 6 ? EpiSpec{1, 1, 1, 1, TC_ACT_NONE, TC_ACT_LRELU}     // Upsample c2 + FiLM1 + residual
                  : EpiSpec{1, 1, 0, 1, TC_ACT_NONE, TC_ACT_NONE};     // Upsample c4 + FiLM2 + residual
}

constexpr int kEpiWarps = 12;      // 4 TMEM lane quarters x 3 column slots
constexpr int kMmaWarp = 12;
constexpr int kProdWarp0 = 13;
constexpr int kMmaWarp2 = 17;     // second MMA-issuing warp: tiles alternate between the two (even: kMmaWarp, odd: kMmaWarp2)

// Walks consecutive tile ids (n_tile * row_tiles + row_tile) without per-tile divisions.
struct TileWalk {
    long long row_tile;
    int n_tile, bq, tt0;     // halo mode: utterance index and first time step of the tile
    int split, jn;           // row-major (top-k) walk: range of channel tiles this sweep covers, position inside it
    __device__ TileWalk(const TcKParams& p, long long tile) {
        split = 0; jn = 0;
        if (p.row_major) {
            // tile id = ((row_tile * splits) + split) * nper + jn: a CTA takes whole (row tile, range) sweeps
            const long long vr = tile / p.topk_nper;
            jn = (int)(tile - vr * p.topk_nper);
            row_tile = vr / p.topk_splits;
            split = (int)(vr - row_tile * p.topk_splits);
            n_tile = split * p.topk_nper + jn;
            bq = 0; tt0 = 0;
            return;
        }
        n_tile = (int)(tile / p.row_tiles);
        row_tile = tile - (long long)n_tile * p.row_tiles;
        bq = p.halo == 1 ? (int)(row_tile / p.tiles_per_utt) : 0;
        tt0 = p.halo == 1 ? (int)(row_tile - (long long)bq * p.tiles_per_utt) * kTileM : 0;
    }
    __device__ void next(const TcKParams& p) {
        if (p.row_major) {
            ++n_tile;
            if (++jn == p.topk_nper) {
                jn = 0;
                if (++split == p.topk_splits) { split = 0; ++row_tile; }
                n_tile = split * p.topk_nper;
            }
            return;
        }
        if (++row_tile == p.row_tiles) { row_tile = 0; ++n_tile; bq = 0; tt0 = 0; return; }
        if (p.halo == 1) {
            tt0 += kTileM;
            if (tt0 >= p.T) { tt0 = 0; ++bq; }
        }
    }
};

// Developer timeline: CTA 0 appends {clock32, role << 28 | event << 24 | (tile & 0xffff) << 8 | (stage & 0xff)} to its
// role's region of p.trace (kTraceRegion entries each).  p.trace == nullptr (always, unless TVC_TC_TRACE is set) costs one
// uniform predicate per site.
constexpr int kTraceRegion = 4096;
struct Tracer {
    uint2* base;
    int n;
    __device__ Tracer(uint2* t, int role) : base((t && blockIdx.x == 0) ? t + role * kTraceRegion : nullptr), n(0) {}
    __device__ __forceinline__ void log(int role, int ev, int tile, int stage) {
        if (base && n < kTraceRegion) {
            base[n++] = make_uint2((uint32_t)clock64(), ((uint32_t)role << 28) | ((uint32_t)ev << 24) | (((uint32_t)tile & 0xffffu) << 8) | ((uint32_t)stage & 0xffu));
        }
    }
    // an event with a time stamp taken earlier (kernel entry), or a value of another clock (globaltimer, ns)
    __device__ __forceinline__ void log_at(int role, int ev, uint32_t stamp) {
        if (base && n < kTraceRegion) base[n++] = make_uint2(stamp, ((uint32_t)role << 28) | ((uint32_t)ev << 24));
    }
};

// ---------------------------------------------------------------------------------------------
// Persistent, warp-specialised: each CTA walks tiles (row tile, channel tile) round-robin.
//   warps 0-11   epilogue: TMEM lane quarter = warp & 3, column slot = warp >> 2 (8-channel groups
//                slot, slot+3, ...); accumulators are double-buffered in TMEM
//   warp  12     TMEM allocation + single-thread tcgen05.mma issue
//   warps 13-16  producers: TMA tensor copies (interior windows, 1x1 / aux stages) or per-thread cp.async gathers (edge
//                windows, flat k = 3 stages) into the smem ring; a producer never waits for its loads
// ---------------------------------------------------------------------------------------------
template <int SPEC>
__global__ void __launch_bounds__(kThreads, 1) tc_conv_kernel(const __grid_constant__ TcKParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    pdl_launch_dependents();      // the next kernel of the plan may start its prologue as soon as SM resources free up
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t t_entry = (uint32_t)clock64();                  // developer timeline only
    const uint32_t stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
    const uint32_t smem_base = smem_u32(smem);
    // barriers: full[ring], empty[ring], acc_full[2], acc_empty[2]
    const uint32_t wreg = smem_base + (uint32_t)p.ring_alloc * stage_bytes;       // resident weight image (w_bytes, may be 0)
    const uint32_t bar_base = wreg + p.w_bytes;
    const uint32_t acc_full = bar_base + 16u * (uint32_t)p.ring;
    const uint32_t acc_empty = acc_full + 16u;
    const uint32_t wfull = acc_full + 40u;                                        // weights-resident: image has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)p.ring_alloc * stage_bytes + p.w_bytes + 16 * p.ring + 32);

    constexpr bool kGeneric = SPEC < 0;
    constexpr EpiSpec kS = epi_spec(SPEC < 0 ? 0 : SPEC);
    const bool film = kGeneric ? p.aux_mode == TC_AUX_FILM : kS.film != 0;
    const int n_main = p.halo ? p.nkb : p.taps * p.nkb;
    const int n_stage = n_main + p.aux_nkb;
    const int chunks = p.KB >> 3;                   // 16-byte chunks per row per plane in one stage
    const uint32_t acc_cols = (uint32_t)(film ? 3 * p.NTp : p.NTp);
    const long long n_tiles_total = p.row_tiles * p.n_tiles;
    // contiguous tile range of this CTA; tile id = n_tile * row_tiles + row_tile
    const long long per_cta = p.row_major ? ((p.row_tiles * p.topk_splits + gridDim.x - 1) / gridDim.x) * p.topk_nper
                                          : (n_tiles_total + gridDim.x - 1) / gridDim.x;
    const long long tile_beg = (long long)blockIdx.x * per_cta;
    const long long tile_end = tile_beg + per_cta < n_tiles_total ? tile_beg + per_cta : n_tiles_total;

    // Set-up is split so that the first loads are in flight ~1 us earlier (profiles/r02h_tc_trace_view.txt: barrier
    // initialisation + TMEM allocation + a CTA-wide barrier took ~2 000 cycles before any producer moved): the lead producer
    // thread initialises the mbarriers, the four producer warps meet on named barrier 2 and start fetching; everybody else
    // (epilogue and MMA warps) additionally waits for the TMEM allocation on named barrier 1, which the first producer warp
    // only signals.
    constexpr int kSetupThreads = kThreads - 4 * 32 + 32;          // non-producers + the signalling producer warp
    uint32_t tmem = 0;
    if (warp >= kProdWarp0 && warp < kMmaWarp2) {
        if (tid == kProdWarp0 * 32) {
            for (int s = 0; s < p.ring; ++s) {
                mbar_init(bar_base + 8u * s, kTileM + 1);          // full: 128 gather arrivals + 1 expect_tx
                mbar_init(bar_base + 8u * (p.ring + s), 1);        // empty: one tcgen05.commit
            }
            mbar_init(wfull, 1);
            mbar_init(acc_full, 1); mbar_init(acc_full + 8, 1);
            mbar_init(acc_empty, kEpiWarps); mbar_init(acc_empty + 8, kEpiWarps);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == kProdWarp0) asm volatile("bar.arrive 1, %0;" ::"n"(kSetupThreads) : "memory");
    } else {
        if (warp == kMmaWarp) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(kSetupThreads) : "memory");
        tc_fence_after();
        tmem = *tmem_slot;
    }

    if (warp >= kProdWarp0 && warp < kMmaWarp2) {
        // ================= producers =================
        // Lane mapping: producer thread j serves tile row j (and window row j + 128 in halo mode) and walks the
        // stage's 8-channel chunks.  With chunk-major operands a warp-level cp.async moves 32 consecutive rows of
        // one chunk: 512 contiguous bytes in global memory and in shared memory (no bank conflicts).
        const int j = tid - kProdWarp0 * 32;
        // bytes between consecutive 8-channel chunk arrays of a plane (gathers; the padded main input never gathers)
        const long long chunk_bytes = p.rows * 16;
        // Ring state per issuing warp: with two MMA warps the ring is split in two halves (tile parity picks the half), so
        // that every thread meets the phases of the barriers it waits on strictly in order (a parity wait cannot tell
        // phase u from phase u - 2).  rs = slot inside the half, rp = parity of the slot's previous use.
        const uint32_t rhalf = p.mma2 ? (uint32_t)p.ring >> 1 : (uint32_t)p.ring;
        uint32_t rs[2] = {0u, 0u}, rp[2] = {0u, 0u};
        bool rwrapped[2] = {false, false};
        const int half = (p.taps - 1) >> 1;
        const uint32_t b_tap = 4u * (uint32_t)p.KB * (uint32_t)p.NTp;
        const uint32_t b_main = p.halo ? (uint32_t)p.taps * b_tap : b_tap;
        const uint32_t b_aux = film ? 2u * b_tap : b_tap;
        TileWalk tw(p, tile_beg);
        // Weights do not depend on the previous kernel: the first tile's weight stages (as many as the ring
        // holds) are requested before griddepcontrol.wait, so they land while the previous grid drains.
        int pre = 0;
        if (p.w_bytes) {
            if (j == 0 && tile_beg < tile_end) {                     // one copy of the whole image (n_tiles == 1)
                mbar_arrive_expect_tx(wfull, p.w_bytes);
                bulk_g2s(wreg, p.w, p.w_bytes, wfull);
            }
        } else if (tile_beg < tile_end && !(p.dbg & 1)) {
            pre = n_stage < (int)rhalf ? n_stage : (int)rhalf;
            if (j == 0) {
                const bf16* w0 = p.w + (long long)tw.n_tile * p.tile_elems;
                for (int i = 0; i < pre; ++i) {
                    const uint32_t bb = i >= n_main ? b_aux : b_main;
                    const uint32_t full = bar_base + 8u * i;
                    mbar_arrive_expect_tx(full, bb);
                    bulk_g2s(smem_base + i * stage_bytes + p.a_stage_bytes, w0, bb, full);
                    w0 += bb >> 1;
                }
            }
        }
        pdl_wait();
        Tracer tr(j == 0 ? p.trace : nullptr, 0);
        for (long long tile = tile_beg; tile < tile_end; ++tile, tw.next(p)) {
            const long long row_tile = tw.row_tile;
            const bf16* wt = p.w + (long long)tw.n_tile * p.tile_elems;
            // flat mode: (utterance base row, time) of the row this thread serves; halo mode: window origin
            int bT = 0, tq = -1;
            int baseT = 0, tt0 = 0;
            if (!p.halo) {
                // rows < 2^31 (checked at the API): 32-bit arithmetic, one division per tile
                const unsigned g = (unsigned)(row_tile * kTileM) + (unsigned)j;
                if ((long long)g < p.rows) {
                    const unsigned bq = g / (unsigned)p.T;
                    bT = (int)(bq * (unsigned)p.T);
                    tq = (int)g - bT;
                }
            } else if (p.halo == 2) {
                // padded mode: tile row j is padded row g of the main input; the aux operand (not padded) is gathered from
                // the row of the same (utterance, time); pad rows produce nothing and get zeros
                const unsigned g = (unsigned)(row_tile * kTileM) + (unsigned)j;
                if ((long long)g < p.rows_a) {
                    const unsigned bq = g / (unsigned)p.Tp;
                    const int t = (int)(g - bq * (unsigned)p.Tp) - p.a_pad;
                    if (t >= 0 && t < p.T) { bT = (int)(bq * (unsigned)p.T); tq = t; }
                }
            } else {
                tt0 = tw.tt0;
                baseT = tw.bq * p.T;
            }
            int tap = 0, kb = 0;
            const int w = p.mma2 ? (int)((tile - tile_beg) & 1) : 0;          // which issuing warp (= ring half) takes this tile
            for (int i = 0; i < n_stage; ++i) {
                const uint32_t s = (uint32_t)w * rhalf + rs[w];
                const uint32_t full = bar_base + 8u * s, empty = bar_base + 8u * (p.ring + s);
                const bool is_aux = i >= n_main;
                const uint32_t b_bytes = is_aux ? b_aux : b_main;
                if (rwrapped[w]) mbar_wait(empty, rp[w]);
                tr.log(0, 0, (int)(tile - tile_beg), i);
                const uint32_t a_dst = smem_base + s * stage_bytes;
                const bf16* src_hi = is_aux ? p.x_hi : p.a_hi;
                const bf16* src_lo = is_aux ? p.x_lo : p.a_lo;
                const int cs = is_aux ? p.x_cs : p.a_cs;
                const uint32_t lbo = is_aux ? p.lbo_aux : p.lbo_main;
                const uint32_t plane = (uint32_t)chunks * lbo;
                // Window geometry.  `start` = operand row that belongs in window slot 0; the window sits `off` = start mod 8
                // slots into the stage so that a TMA box (which starts on a multiple of 8 rows) lands it in place; the MMA
                // warp applies the same offset to its descriptors.  Tiles whose window leaves the utterance (replicate
                // padding) and k = 3 stages of flat mode are gathered per thread with cp.async instead.
                int win = kTileM, org = 0;
                long long start;
                bool edge = false;
                if (p.halo == 1) {
                    win = is_aux ? kTileM : p.R;
                    org = tt0 - (is_aux ? 0 : p.dil);
                    start = (long long)baseT + org;
                    edge = org < 0 || org + win > p.T;
                } else if (p.halo == 2) {
                    start = row_tile * kTileM - (is_aux ? 0 : p.dil);      // padded rows: the window never leaves its utterance's pads
                } else {
                    start = row_tile * kTileM;
                }
                const uint32_t off = (uint32_t)(start + 64) & 7u;
                const bool tma = (is_aux ? p.tma_aux != 0 : (p.tma_main != 0 && (p.halo || p.taps == 1))) && !edge && !(p.dbg & 1);
                // issue work is spread over the first lanes of three producer warps (a bulk / tensor copy costs its issuing
                // thread a few hundred cycles): lane 0 of warp 0 books the bytes and fetches the weights, warps 1 and 2 the planes
                if (j == 0) {
                    const uint32_t t_bytes = tma ? 2u * plane : 0u;
                    if (tile == tile_beg && i < pre) { if (tma) mbar_expect_tx(full, t_bytes); }       // weights already requested
                    else if (p.dbg & 1) mbar_arrive(full);
                    else if (p.w_bytes) mbar_arrive_expect_tx(full, t_bytes);                           // weights are resident
                    else {
                        mbar_arrive_expect_tx(full, b_bytes + t_bytes);
                        bulk_g2s(a_dst + p.a_stage_bytes, wt, b_bytes, full);
                    }
                } else if (tma && j == 32) {
                    tma_load_3d(a_dst, is_aux ? &p.tm_x_hi : &p.tm_a_hi, 0, (int)((start - off) >> 3), kb * chunks, full);
                } else if (tma && j == 64) {
                    tma_load_3d(a_dst + plane, is_aux ? &p.tm_x_lo : &p.tm_a_lo, 0, (int)((start - off) >> 3), kb * chunks, full);
                }
                wt += b_bytes >> 1;
                int n_ok = (cs >> 3) - kb * chunks;            // chunks of this stage that exist in the tensor; the rest are zero-filled
                n_ok = n_ok > chunks ? chunks : n_ok;
                // Pointer-walking gathers: per 16-byte copy only a 64-bit pointer bump and a shared-address bump remain
                const char* g_hi = reinterpret_cast<const char*>(src_hi) + (long long)kb * chunks * chunk_bytes;
                const char* g_lo = reinterpret_cast<const char*>(src_lo) + (long long)kb * chunks * chunk_bytes;
                if ((p.dbg & 1) || tma) {
                } else if (p.halo != 1) {
                    const int shift = is_aux ? 0 : (tap - half) * p.dil;
                    int tt = tq + shift;
                    tt = tt < 0 ? 0 : (tt > p.T - 1 ? p.T - 1 : tt);                    // replicate padding
                    const long long rb = (long long)(bT + tt) * 16;                     // byte offset of the row inside a chunk array
                    const uint32_t sz = tq >= 0 ? 16u : 0u;                             // 0 -> zero fill (rows past the end)
                    const char* ph = g_hi + rb;
                    const char* pl = g_lo + rb;
                    uint32_t dst = a_dst + (uint32_t)j * 16u;                           // flat tiles start on a multiple of 128 rows: off = 0
                    int c = 0;
#pragma unroll 4
                    for (; c < n_ok; ++c) {
                        cp_async16(dst, ph, sz);
                        cp_async16(dst + plane, pl, sz);
                        ph += chunk_bytes; pl += chunk_bytes; dst += lbo;
                    }
                    for (; c < chunks; ++c) {                                           // channel padding of the K-stage
                        cp_async16(dst, g_hi, 0u);
                        cp_async16(dst + plane, g_lo, 0u);
                        dst += lbo;
                    }
                } else {
                    for (int m = j; m < win; m += kTileM) {
                        int tt = org + m;
                        tt = tt < 0 ? 0 : (tt > p.T - 1 ? p.T - 1 : tt);                // replicate padding
                        const long long rb = (long long)(baseT + tt) * 16;
                        const char* ph = g_hi + rb;
                        const char* pl = g_lo + rb;
                        uint32_t dst = a_dst + (off + (uint32_t)m) * 16u;
                        int c = 0;
#pragma unroll 4
                        for (; c < n_ok; ++c) {
                            cp_async16(dst, ph, 16u);
                            cp_async16(dst + plane, pl, 16u);
                            ph += chunk_bytes; pl += chunk_bytes; dst += lbo;
                        }
                        for (; c < chunks; ++c) {
                            cp_async16(dst, g_hi, 0u);
                            cp_async16(dst + plane, g_lo, 0u);
                            dst += lbo;
                        }
                    }
                }
                cp_async_arrive_noinc(full);
                tr.log(0, 1, (int)(tile - tile_beg), i);
                // stage order: K-block major, tap minor (flat); K-block only (halo); then the aux K-blocks
                if (is_aux || p.halo) { ++kb; }
                else if (++tap == p.taps) { tap = 0; ++kb; }
                if (i + 1 == n_main) { kb = 0; tap = 0; }
                if (++rs[w] == rhalf) { rs[w] = 0; rp[w] = rwrapped[w] ? rp[w] ^ 1u : 0u; rwrapped[w] = true; }
            }
        }
    } else if (warp == kMmaWarp || warp == kMmaWarp2) {
        // ================= MMA issuers (one elected lane; the warp runs the loops uniformly) =================
        // This warp is a single instruction stream and, once the operands arrive by TMA, the kernel's critical path: a
        // full-rate 24-channel tile is 18 MMAs (~940 cycles of blocking issue) and used to cost as much again in scalar
        // work between them.  So: kernel parameters live in registers (not re-read from the constant bank), the tile
        // walk is a few adds, main and aux stages have their own loops with pre-built descriptor words, and a stage's
        // TAPS x KS x 3 MMAs are straight-line code.
        {
            const bool leader = elect_one() != 0;
            const bool issue = leader && !(p.dbg & 2);
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);      // SBO = 128 bytes, descriptor version 1
            const int ksteps = p.KB >> 4;
            uint32_t ring = (uint32_t)p.ring, NTp = (uint32_t)p.NTp, dil = (uint32_t)p.dil;
            uint32_t stage16 = stage_bytes >> 4;
            const uint32_t base16 = (smem_base & 0x3FFFFu) >> 4;
            // low descriptor words without the moving start address: LBO field, and for B the fixed offset inside a stage
            uint32_t lbo_m16 = p.lbo_main >> 4, lbo_x16 = p.lbo_aux >> 4;
            uint32_t plane_m16 = (uint32_t)chunks * lbo_m16, plane_x16 = (uint32_t)chunks * lbo_x16;
            uint32_t a_fld_m = lbo_m16 << 16, a_fld_x = lbo_x16 << 16;
            const uint32_t lbo_bx = film ? 2u * NTp : NTp;                 // weight-image chunk stride: rows of the image
            uint32_t plane_bm = (uint32_t)chunks * NTp, plane_bx = (uint32_t)chunks * lbo_bx;
            // B descriptor words: inside the stage (after the A window), or inside the resident image (stage i at a fixed offset)
            const bool wres = p.w_bytes != 0;
            const uint32_t bimg16 = wres ? (wreg & 0x3FFFFu) >> 4 : base16 + (p.a_stage_bytes >> 4);
            uint32_t b_fld_m = bimg16 | (NTp << 16), b_fld_x = bimg16 | (lbo_bx << 16);
            uint32_t bstep_m = 0, bstep_x = 0, bx0 = 0;                  // weights-resident: image offsets of the stages (16-byte units)
            if (wres) {
                const uint32_t b_tap16 = (4u * (uint32_t)p.KB * NTp) >> 4;
                bstep_m = p.halo ? (uint32_t)p.taps * b_tap16 : b_tap16;
                bstep_x = film ? 2u * b_tap16 : b_tap16;
                bx0 = (uint32_t)n_main * bstep_m;
            }
            uint32_t idesc_m = umma_idesc(kTileM, NTp), idesc_x = film ? umma_idesc(kTileM, 2u * NTp) : idesc_m;
            const bool halo_main = p.halo != 0;
            const bool utt_walk = p.halo == 1;       // padded mode: tiles start on multiples of 128 padded rows (row0 stays 0 mod 8)
            int n_aux = p.aux_nkb;
            // tile walk state (halo mode: operand row of the tile's first output row; only its low 3 bits matter)
            uint32_t T = (uint32_t)p.T, tt0 = 0, row0 = 0;
            long long rt_left = 0;
            {
                TileWalk tw(p, tile_beg);
                tt0 = (uint32_t)tw.tt0;
                row0 = (uint32_t)((long long)tw.bq * p.T + tw.tt0);
                rt_left = p.row_tiles - tw.row_tile;                      // row tiles until the walk wraps to the next channel tile
            }
            // keep them in registers: opaque to the optimiser, so no re-materialisation through LDC in the loops
            asm volatile("" : "+r"(ring), "+r"(NTp), "+r"(dil), "+r"(stage16), "+r"(lbo_m16), "+r"(lbo_x16), "+r"(plane_m16), "+r"(plane_x16));
            asm volatile("" : "+r"(a_fld_m), "+r"(a_fld_x), "+r"(plane_bm), "+r"(plane_bx), "+r"(b_fld_m), "+r"(b_fld_x), "+r"(idesc_m), "+r"(idesc_x));
            asm volatile("" : "+r"(T), "+r"(n_aux));
            uint32_t s = 0, ph = 0, tcount = 0;
            // Two issuing warps alternate tiles (tile parity = TMEM buffer = half of the smem ring), so one warp's per-tile
            // scalar work (barrier waits, descriptor set-up, commits) overlaps the other's blocking MMA issue.  Each warp owns
            // its half of the ring: sharing one ring let a parity wait fall through two uses early (an mbarrier parity wait
            // cannot tell phase u from phase u - 2 unless the thread has passed every phase in order).
            const uint32_t mine = warp == kMmaWarp ? 0u : 1u;
            const uint32_t rhalf = p.mma2 ? (uint32_t)p.ring >> 1 : (uint32_t)p.ring;
            const uint32_t sbase = mine * rhalf;                          // first slot of this warp's half (single issuer: 0)
            ring = rhalf;                                                  // wrap point of this warp's slot counter
            Tracer tr((leader && mine == 0) ? p.trace : nullptr, 1);
            if (p.trace) {                                                 // ev 6: kernel entry, ev 8: globaltimer (ns) now
                unsigned long long gt;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
                tr.log_at(1, 6, t_entry);
                tr.log_at(1, 8, (uint32_t)gt);
                tr.log(1, 7, 0, 0);                                        // ev 7: the MMA warp is ready to issue
            }
            const long long n_my = tile_end - tile_beg;
            if (wres && n_my > 0) { mbar_wait(wfull, 0u); tc_fence_after(); }
            for (long long it = 0; it < n_my; ++it, ++tcount) {
              if (p.mma2 ? (tcount & 1u) != mine : mine != 0u) {
                // the other issuer's tile (single-issuer mode: the second warp idles)
              } else {
                const uint32_t buf = tcount & 1u, buse = tcount >> 1;
                tr.log(1, 4, (int)tcount, 0);
                if (buse > 0) {                                         // epilogue must have drained this accumulator
                    mbar_wait(acc_empty + 8u * buf, (buse - 1) & 1u);
                    tc_fence_after();
                }
                tr.log(1, 0, (int)tcount, 0);
                // where window slot 0 sits inside a stage (see the producers): operand row of slot 0, modulo 8
                const uint32_t off_m = halo_main ? ((row0 - dil + 64u) & 7u) : 0u;
                const uint32_t off_x = halo_main ? (row0 & 7u) : 0u;
                const uint32_t d_base = tmem + buf * acc_cols;
#define TVC_ISSUE(T_, K_, A_LO, B_LO, PA, PB, LBA, LBB, IDESC, D_, ACC0) \
    issue_stage<T_, K_>(D_, A_LO, B_LO, dil, 2u * (PB), 2u * (LBA), 2u * (LBB), PA, PB, desc_hi, IDESC, ACC0)
                // ---- main stages
                for (int i = 0; i < n_main; ++i) {
                    const uint32_t full = bar_base + 8u * (sbase + s), empty = bar_base + 8u * ((uint32_t)p.ring + sbase + s);
                    const uint32_t so = (sbase + s) * stage16;
                    const uint32_t a_lo0 = (a_fld_m + base16 + so + off_m), b_lo0 = b_fld_m + (wres ? (uint32_t)i * bstep_m : so);
                    const uint32_t acc0 = i == 0 ? 0u : 1u;
                    tr.log(1, 5, (int)tcount, i);
                    mbar_wait(full, ph);
                    tc_fence_after();
                    tr.log(1, 1, (int)tcount, i);
                    if (issue) {
                        // halo mode: a tap is a row offset into the shared window (one 16-byte slot per row) and
                        // selects the tap's weight image (hi + lo planes apart)
                        if (halo_main) {
                            switch (ksteps) {
                                case 1: TVC_ISSUE(3, 1, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                                case 2: TVC_ISSUE(3, 2, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                                case 3: TVC_ISSUE(3, 3, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                                default: TVC_ISSUE(3, 4, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                            }
                        } else {
                            switch (ksteps) {
                                case 1: TVC_ISSUE(1, 1, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                                case 2: TVC_ISSUE(1, 2, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                                case 3: TVC_ISSUE(1, 3, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                                default: TVC_ISSUE(1, 4, a_lo0, b_lo0, plane_m16, plane_bm, lbo_m16, NTp, idesc_m, d_base, acc0); break;
                            }
                        }
                    }
                    if (leader) umma_commit(empty);                    // frees the smem stage once these MMAs retire
                    tr.log(1, 2, (int)tcount, i);
                    if (++s == ring) { s = 0; ph ^= 1u; }
                }
                // ---- aux stages (1x1 on the second input): accumulate into the same columns (TC_AUX_ACC) or feed the
                //      FiLM scale|shift accumulator next to them
                for (int i = 0; i < n_aux; ++i) {
                    const uint32_t full = bar_base + 8u * (sbase + s), empty = bar_base + 8u * ((uint32_t)p.ring + sbase + s);
                    const uint32_t so = (sbase + s) * stage16;
                    const uint32_t a_lo0 = (a_fld_x + base16 + so + off_x), b_lo0 = b_fld_x + (wres ? bx0 + (uint32_t)i * bstep_x : so);
                    const uint32_t d = d_base + (film ? NTp : 0u);
                    const uint32_t acc0 = (film && i == 0) ? 0u : 1u;
                    tr.log(1, 5, (int)tcount, n_main + i);
                    mbar_wait(full, ph);
                    tc_fence_after();
                    tr.log(1, 1, (int)tcount, n_main + i);
                    if (issue) {
                        switch (ksteps) {
                            case 1: TVC_ISSUE(1, 1, a_lo0, b_lo0, plane_x16, plane_bx, lbo_x16, lbo_bx, idesc_x, d, acc0); break;
                            case 2: TVC_ISSUE(1, 2, a_lo0, b_lo0, plane_x16, plane_bx, lbo_x16, lbo_bx, idesc_x, d, acc0); break;
                            case 3: TVC_ISSUE(1, 3, a_lo0, b_lo0, plane_x16, plane_bx, lbo_x16, lbo_bx, idesc_x, d, acc0); break;
                            default: TVC_ISSUE(1, 4, a_lo0, b_lo0, plane_x16, plane_bx, lbo_x16, lbo_bx, idesc_x, d, acc0); break;
                        }
                    }
                    if (leader) umma_commit(empty);
                    tr.log(1, 2, (int)tcount, n_main + i);
                    if (++s == ring) { s = 0; ph ^= 1u; }
                }
#undef TVC_ISSUE
                if (leader) umma_commit(acc_full + 8u * buf);          // accumulators complete -> epilogue
                tr.log(1, 3, (int)tcount, 0);
              }
                // next tile of the walk (row tiles first, then the next channel tile starts again at row 0)
                if (--rt_left == 0) { rt_left = p.row_tiles; tt0 = 0; row0 = 0; }
                else if (utt_walk) {
                    tt0 += kTileM; row0 += kTileM;
                    if (tt0 >= T) { row0 += T - tt0; tt0 = 0; }        // first row of the next utterance
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue =================
        const bool has_res = kGeneric ? p.res != nullptr : kS.res != 0;
        const bool has_y32 = kGeneric ? p.y32 != nullptr : kS.y32 != 0;
        const bool has_pl = kGeneric ? p.y_hi != nullptr : kS.planes != 0;
        const int epi_act = kGeneric ? p.epi_act : kS.epi_act;
        const int out_act = kGeneric ? p.out_act : kS.out_act;
        const int quarter = warp & 3, slot = warp >> 2;
        const int rloc = quarter * 32 + lane;
        pdl_wait();               // residual reads and output writes must follow the previous grid
        const int n_groups = p.NT >> 3;
        uint32_t tcount = 0;
        TileWalk tw(p, tile_beg);
        // fused top-k state (SPEC 8): this thread's best columns of its row so far, best first
        [[maybe_unused]] float tkv[kTopkC];
        [[maybe_unused]] int tki[kTopkC];
        const int trole = warp == 0 ? 2 : 3;
        Tracer tr((lane == 0 && (warp == 0 || warp == kEpiWarps - 1)) ? p.trace : nullptr, trole);
        for (long long tile = tile_beg; tile < tile_end; ++tile, ++tcount, tw.next(p)) {
            const uint32_t buf = tcount & 1u, buse = tcount >> 1;
            const int n_tile = tw.n_tile;
            tr.log(trole, 0, (int)tcount, 0);
            long long row, rowp;                 // output row (residual, fp32 output) and row of the plane output
            bool valid, first = false, last = false;
            [[maybe_unused]] int ut = 0;         // SPEC 9: time step inside the utterance and the utterance index
            [[maybe_unused]] long long ub = 0;
            if (!p.halo) {
                row = tw.row_tile * kTileM + rloc;
                valid = row < p.rows;
                rowp = row;
                if constexpr (!kGeneric && kS.dec != 0) {
                    ub = (long long)((unsigned)row / (unsigned)p.T);
                    ut = (int)(row - ub * p.T);
                }
            } else if (p.halo == 2) {
                // padded mode: tile rows are padded input rows; pad rows produce nothing.  The plane output may carry its own
                // padding (y_pad rows on either side of every utterance): the threads that own t = 0 and t = T - 1 also store
                // the replicate rows, so the next conv's windows never leave the tensor (replicate padding, decoder.py:144-146).
                const unsigned g = (unsigned)(tw.row_tile * kTileM) + (unsigned)rloc;
                valid = false;
                row = 0; rowp = 0;
                if ((long long)g < p.rows_a) {
                    const unsigned bq = g / (unsigned)p.Tp;
                    const int t = (int)(g - bq * (unsigned)p.Tp) - p.a_pad;
                    if (t >= 0 && t < p.T) {
                        valid = true;
                        ut = t; ub = bq;
                        row = (long long)bq * p.T + t;
                        rowp = (long long)bq * (p.T + 2 * p.y_pad) + p.y_pad + t;
                        first = t == 0;
                        last = t == p.T - 1;
                    }
                }
            } else {
                const int t = tw.tt0 + rloc;
                row = (long long)tw.bq * p.T + t;
                valid = t < p.T;
                rowp = row;
                ut = t; ub = tw.bq;
            }
            const uint32_t lane_addr = tmem + buf * acc_cols + ((uint32_t)(quarter * 32) << 16);
            const float* bias = p.bias + n_tile * p.NTp;               // padded channel space
            const float* fbias = film ? p.film_bias + 2 * n_tile * p.NTp : nullptr;
            const int ch0 = n_tile * p.NT;                             // first real output channel of this tile
            if constexpr (!kGeneric && kS.topk != 0) {
                // ---- kNN screening (feature_retrieval.py:27-28): the similarity tile never leaves the SM.  A thread sees the
                //      columns of its row in increasing order (its column groups inside a tile, the tiles of the sweep), so a
                //      strict '>' keeps the lower index on ties, like torch.topk on the CPU.
                if (tw.jn == 0) {
#pragma unroll
                    for (int i = 0; i < kTopkC; ++i) { tkv[i] = -INFINITY; tki[i] = -1; }
                }
                mbar_wait(acc_full + 8u * buf, buse & 1u);
                tc_fence_after();
                for (int cg = slot; cg < n_groups; cg += 3) {
                    float v[8];
                    tmem_ld8(lane_addr + (uint32_t)(cg * 8), v);
                    tmem_ld_wait();
                    const int n0 = ch0 + cg * 8;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        if (v[e] > tkv[kTopkC - 1] && n0 + e < p.topk_n) {
                            tkv[kTopkC - 1] = v[e];                  // replace the worst, then bubble up (static indices only:
                            tki[kTopkC - 1] = n0 + e;                // the lists must stay in registers)
#pragma unroll
                            for (int q = kTopkC - 1; q > 0; --q)
                                if (tkv[q] > tkv[q - 1]) {
                                    const float tv = tkv[q]; tkv[q] = tkv[q - 1]; tkv[q - 1] = tv;
                                    const int ti = tki[q]; tki[q] = tki[q - 1]; tki[q - 1] = ti;
                                }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty + 8u * buf);
                if (tw.jn == p.topk_nper - 1) {
                    // end of the row tile's sweep: the three column slots of a lane quarter merge their lists through shared memory
                    float* mv = reinterpret_cast<float*>(smem + p.topk_smem);            // [3 slots][128 rows][kTopkC]
                    int* mi = reinterpret_cast<int*>(mv + 3 * kTileM * kTopkC);
#pragma unroll
                    for (int i = 0; i < kTopkC; ++i) {
                        mv[(slot * kTileM + rloc) * kTopkC + i] = tkv[i];
                        mi[(slot * kTileM + rloc) * kTopkC + i] = tki[i];
                    }
                    asm volatile("bar.sync %0, 96;" ::"r"(3 + quarter) : "memory");
                    if (slot == 0 && valid) {
                        // three-way merge of the sorted lists, straight from shared memory (cold path: scalars only, so
                        // that the hot loop's lists are the only long-lived registers)
                        int h0 = 0, h1 = 0, h2 = 0;
                        float vk = 0.f, v8 = -INFINITY;
                        for (int i = 0; i < kTopkC; ++i) {
                            const int e0 = rloc * kTopkC + h0, e1 = (kTileM + rloc) * kTopkC + h1, e2 = (2 * kTileM + rloc) * kTopkC + h2;
                            const int n0 = h0 < kTopkC ? mi[e0] : -1, n1 = h1 < kTopkC ? mi[e1] : -1, n2 = h2 < kTopkC ? mi[e2] : -1;
                            const float x0 = n0 >= 0 ? mv[e0] : -INFINITY, x1 = n1 >= 0 ? mv[e1] : -INFINITY, x2 = n2 >= 0 ? mv[e2] : -INFINITY;
                            int w = -1, bn = -1;
                            float bx = -INFINITY;
                            if (n0 >= 0) { w = 0; bn = n0; bx = x0; }
                            if (n1 >= 0 && (w < 0 || x1 > bx || (x1 == bx && n1 < bn))) { w = 1; bn = n1; bx = x1; }
                            if (n2 >= 0 && (w < 0 || x2 > bx || (x2 == bx && n2 < bn))) { w = 2; bn = n2; bx = x2; }
                            h0 += w == 0; h1 += w == 1; h2 += w == 2;
                            const long long o = (row * p.topk_splits + tw.split) * kTopkC + i;
                            p.topk_cand[o] = bn;
                            if (p.topk_score) p.topk_score[o] = bx;
                            if (i == p.topk_k - 1) vk = bx;
                            if (i == kTopkC - 1) v8 = bx;
                        }
                        // fewer columns than candidates: every column is a candidate, nothing can be missed
                        const bool thin = p.topk_n > kTopkC && !(vk - v8 > p.topk_eps);      // also true for NaN
                        if (p.topk_splits == 1) p.topk_flag[row] = thin ? 1 : 0;            // split sweep: decided after the merge
                    }
                    asm volatile("bar.sync %0, 96;" ::"r"(3 + quarter) : "memory");          // scratch may be rewritten
                }
                tr.log(trole, 3, (int)tcount, 0);
                continue;
            }
            bool waited = false;
            for (int cg = slot; cg < n_groups; cg += 3) {
                const int ch = ch0 + cg * 8;
                const bool live = valid && ch < p.Cout && !(p.dbg & 4);
                float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
                if (has_res && live) {                                 // issued before the accumulator wait
                    const float4* rp = reinterpret_cast<const float4*>(p.res + ((long long)(ch >> 3) * p.rows + row) * 8);
                    r0 = __ldg(rp);
                    r1 = __ldg(rp + 1);
                }
                // per-channel constants: requested before the accumulator wait so their latency hides behind it
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cg * 8) + 1);
                float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, h0 = s0, h1 = s0;
                if (film) {
                    s0 = __ldg(reinterpret_cast<const float4*>(fbias + cg * 8));
                    s1 = __ldg(reinterpret_cast<const float4*>(fbias + cg * 8) + 1);
                    h0 = __ldg(reinterpret_cast<const float4*>(fbias + p.NTp + cg * 8));
                    h1 = __ldg(reinterpret_cast<const float4*>(fbias + p.NTp + cg * 8) + 1);
                }
                if (!waited) {
                    mbar_wait(acc_full + 8u * buf, buse & 1u);
                    tc_fence_after();
                    waited = true;
                    tr.log(trole, 1, (int)tcount, 0);
                }
                float v[8], sc[8], sh[8];
                tmem_ld8(lane_addr + (uint32_t)(cg * 8), v);
                if (film) {
                    tmem_ld8(lane_addr + (uint32_t)(p.NTp + cg * 8), sc);
                    tmem_ld8(lane_addr + (uint32_t)(2 * p.NTp + cg * 8), sh);
                }
                tmem_ld_wait();
                tr.log(trole, 2, (int)tcount, cg);
                {
                    v[0] = __fadd_rn(v[0], b0.x); v[1] = __fadd_rn(v[1], b0.y); v[2] = __fadd_rn(v[2], b0.z); v[3] = __fadd_rn(v[3], b0.w);
                    v[4] = __fadd_rn(v[4], b1.x); v[5] = __fadd_rn(v[5], b1.y); v[6] = __fadd_rn(v[6], b1.z); v[7] = __fadd_rn(v[7], b1.w);
                }
                [[maybe_unused]] float vn[8];            // SPEC 9: the same channels of the next time row (the resampler's second tap)
                if constexpr (!kGeneric && kS.dec != 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float nx = __shfl_down_sync(0xffffffffu, v[i], 1);
                        vn[i] = lane == 31 ? v[i] : nx;  // only reached with weight 0 (factors 5 and 3 sample one row exactly)
                    }
                }
                if (!live) continue;
                if (film) {
                    const float sb[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    const float hb[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i)                        // FiLM: x * scale + shift (decoder.py:97)
                        v[i] = __fadd_rn(__fmul_rn(v[i], __fadd_rn(sc[i], sb[i])), __fadd_rn(sh[i], hb[i]));
                }
                if (has_res) {
                    v[0] = __fadd_rn(v[0], r0.x); v[1] = __fadd_rn(v[1], r0.y); v[2] = __fadd_rn(v[2], r0.z); v[3] = __fadd_rn(v[3], r0.w);
                    v[4] = __fadd_rn(v[4], r1.x); v[5] = __fadd_rn(v[5], r1.y); v[6] = __fadd_rn(v[6], r1.z); v[7] = __fadd_rn(v[7], r1.w);
                }
                if (epi_act != TC_ACT_NONE) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i], epi_act);
                }
                const long long o = ((long long)(ch >> 3) * p.rows + row) * 8;     // chunk-major: (row, chunk) -> 8 elements
                if (has_y32 && ch + 8 <= p.y32_cs) {
                    reinterpret_cast<float4*>(p.y32 + o)[0] = make_float4(v[0], v[1], v[2], v[3]);
                    reinterpret_cast<float4*>(p.y32 + o)[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
                if (has_pl && ch + 8 <= p.y_cs) {
                    uint32_t h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) split2(apply_act(v[2 * i], out_act), apply_act(v[2 * i + 1], out_act), h[i], l[i]);
                    const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(l[0], l[1], l[2], l[3]);
                    const long long op = ((long long)(ch >> 3) * p.rows_y + rowp) * 8;
                    *reinterpret_cast<uint4*>(p.y_hi + op) = hv;
                    *reinterpret_cast<uint4*>(p.y_lo + op) = lv;
                    if (first) {
                        for (int q = 1; q <= p.y_pad; ++q) {
                            *reinterpret_cast<uint4*>(p.y_hi + op - 8 * q) = hv;
                            *reinterpret_cast<uint4*>(p.y_lo + op - 8 * q) = lv;
                        }
                    }
                    if (last) {
                        for (int q = 1; q <= p.y_pad; ++q) {
                            *reinterpret_cast<uint4*>(p.y_hi + op + 8 * q) = hv;
                            *reinterpret_cast<uint4*>(p.y_lo + op + 8 * q) = lv;
                        }
                    }
                }
                if constexpr (!kGeneric && kS.dec != 0) {
                    // F.interpolate(scale_factor = 1 / f) of this conv's output, interp_cl's arithmetic: output d of the
                    // utterance reads rows i0(d), i1(d); the thread that owns row i0 produces it (its neighbour lane owns i1).
                    const int Tout = p.T / p.dec_f;
                    const int d = ut / p.dec_f;
                    if (d < Tout && ch + 8 <= p.y_cs) {
                        const LinCoord c = lin_coord(d, p.dec_scale, p.T);
                        if (c.i0 == ut) {
                            float w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) w[i] = lin_blend(v[i], c.i1 == ut ? v[i] : vn[i], c);
                            const long long rows_r = (p.rows / p.T) * Tout;
                            uint32_t h[4], l[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) split2(w[2 * i], w[2 * i + 1], h[i], l[i]);
                            const long long orr = ((long long)(ch >> 3) * rows_r + ub * Tout + d) * 8;
                            *reinterpret_cast<uint4*>(p.dec_r_hi + orr) = make_uint4(h[0], h[1], h[2], h[3]);
                            *reinterpret_cast<uint4*>(p.dec_r_lo + orr) = make_uint4(l[0], l[1], l[2], l[3]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) split2(leaky01(w[2 * i]), leaky01(w[2 * i + 1]), h[i], l[i]);
                            const uint4 hv = make_uint4(h[0], h[1], h[2], h[3]), lv = make_uint4(l[0], l[1], l[2], l[3]);
                            const long long Tpo = Tout + 2 * p.dec_pad;
                            const long long oa = ((long long)(ch >> 3) * ((p.rows / p.T) * Tpo) + ub * Tpo + p.dec_pad + d) * 8;
                            *reinterpret_cast<uint4*>(p.dec_a_hi + oa) = hv;
                            *reinterpret_cast<uint4*>(p.dec_a_lo + oa) = lv;
                            if (d == 0)
                                for (int q = 1; q <= p.dec_pad; ++q) {
                                    *reinterpret_cast<uint4*>(p.dec_a_hi + oa - 8 * q) = hv;
                                    *reinterpret_cast<uint4*>(p.dec_a_lo + oa - 8 * q) = lv;
                                }
                            if (d == Tout - 1)
                                for (int q = 1; q <= p.dec_pad; ++q) {
                                    *reinterpret_cast<uint4*>(p.dec_a_hi + oa + 8 * q) = hv;
                                    *reinterpret_cast<uint4*>(p.dec_a_lo + oa + 8 * q) = lv;
                                }
                        }
                    }
                }
            }
            if (!waited) {                                             // a slot without column groups still takes part
                mbar_wait(acc_full + 8u * buf, buse & 1u);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + 8u * buf);          // accumulator buffer may be overwritten
            tr.log(trole, 3, (int)tcount, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.trace && blockIdx.x == 0 && tid == 0) {                  // developer timeline: all roles of CTA 0 are done (ev 9, ev 10 = ns)
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        p.trace[4 * kTraceRegion - 2] = make_uint2((uint32_t)clock64(), (3u << 28) | (9u << 24));
        p.trace[4 * kTraceRegion - 1] = make_uint2((uint32_t)gt, (3u << 28) | (10u << 24));
    }
    if (warp == kMmaWarp) tmem_dealloc(tmem, p.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------
inline uint16_t f2bf_host(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0;
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf2f_host(uint16_t h) {
    const uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

int pick_kb(int cin) {
    const int c16 = (int)align_up(cin, 16);
    if (c16 <= 64) return c16;
    if (c16 % 64 == 0) return 64;
    if (c16 % 48 == 0) return 48;
    return 64;   // K is zero-padded up to a multiple of 64
}

}  // namespace

void TcConvW::free_all() {
    if (w) cudaFree(w);
    if (bias) cudaFree(bias);
    if (film_bias) cudaFree(film_bias);
    w = nullptr; bias = nullptr; film_bias = nullptr;
}

int tc_pack_conv(const float* w, const float* b, int Cout, int Cin, int taps, const float* aux_w, const float* aux_b,
                 int aux_cin, int aux_mode, int NT, TcConvW& o, bool cat) {
    TVC_REQUIRE(Cout > 0 && Cin > 0 && (taps == 1 || taps == 3), "tc_pack_conv: bad shape Cout=%d Cin=%d taps=%d", Cout, Cin, taps);
    TVC_REQUIRE(NT % 8 == 0 && NT >= 8, "tc_pack_conv: NT=%d must be a multiple of 8", NT);
    o.Cin = Cin; o.Cout = Cout; o.taps = taps; o.aux_cin = aux_mode ? aux_cin : 0; o.aux_mode = aux_mode;
    o.KB = pick_kb(aux_mode ? (Cin > aux_cin ? Cin : aux_cin) : Cin);
    o.nkb = cdiv(align_up(Cin, 16), o.KB);
    o.aux_nkb = aux_mode ? cdiv(align_up(aux_cin, 16), o.KB) : 0;
    o.NT = NT; o.NTp = (int)align_up(NT, 16); o.n_tiles = cdiv(Cout, NT);
    const int film_rows = 2 * o.NTp;
    TVC_REQUIRE(o.NTp <= 256 && (aux_mode != TC_AUX_FILM || film_rows <= 256), "tc_pack_conv: NT=%d too wide", NT);
    const int chunks = o.KB / 8;
    const size_t main_stage = (size_t)2 * chunks * o.NTp * 8;                                   // bf16 elements, hi+lo
    const size_t aux_stage = (size_t)2 * chunks * (aux_mode == TC_AUX_FILM ? film_rows : o.NTp) * 8;
    o.tile_elems = main_stage * taps * o.nkb + aux_stage * o.aux_nkb;
    std::vector<uint16_t> img(o.tile_elems * o.n_tiles, 0);
    std::vector<float> hb((size_t)o.n_tiles * o.NTp, 0.f), hf;
    if (aux_mode == TC_AUX_FILM) hf.assign((size_t)o.n_tiles * film_rows, 0.f);

    o.cat = cat;
    auto put = [&](uint16_t* stage, int n_rows, int c, int n, int e, float v) {
        const uint16_t h = f2bf_host(v);
        const uint16_t l = f2bf_host(v - bf2f_host(h));
        const size_t plane = (size_t)chunks * n_rows * 8;
        const size_t off = ((size_t)c * n_rows + n) * 8 + e;
        stage[off] = h;
        stage[plane + off] = l;
    };
    // main stages of a "cat" image: [chunk][hi rows | lo rows][8]
    auto put_cat = [&](uint16_t* stage, int c, int n, int e, float v) {
        const uint16_t h = f2bf_host(v);
        const uint16_t l = f2bf_host(v - bf2f_host(h));
        const size_t off = ((size_t)c * 2 * o.NTp + n) * 8 + e;
        stage[off] = h;
        stage[off + (size_t)o.NTp * 8] = l;
    };
    for (int nt = 0; nt < o.n_tiles; ++nt) {
        uint16_t* tile = img.data() + o.tile_elems * nt;
        size_t so = 0;
        for (int kb = 0; kb < o.nkb; ++kb)
            for (int tap = 0; tap < taps; ++tap, so += main_stage)
                for (int c = 0; c < chunks; ++c)
                    for (int n = 0; n < NT; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int co = nt * NT + n, ci = kb * o.KB + c * 8 + e;
                            if (co < Cout && ci < Cin) {
                                const float wv = w[((size_t)co * Cin + ci) * taps + tap];
                                if (o.cat) put_cat(tile + so, c, n, e, wv);
                                else put(tile + so, o.NTp, c, n, e, wv);
                            }
                        }
        for (int kb = 0; kb < o.aux_nkb; ++kb, so += aux_stage)
            for (int c = 0; c < chunks; ++c)
                for (int n = 0; n < NT; ++n)
                    for (int e = 0; e < 8; ++e) {
                        const int co = nt * NT + n, ci = kb * o.KB + c * 8 + e;
                        if (co >= Cout || ci >= aux_cin) continue;
                        if (aux_mode == TC_AUX_ACC) {
                            put(tile + so, o.NTp, c, n, e, aux_w[(size_t)co * aux_cin + ci]);
                        } else {
                            put(tile + so, film_rows, c, n, e, aux_w[(size_t)co * aux_cin + ci]);                          // scale
                            put(tile + so, film_rows, c, o.NTp + n, e, aux_w[((size_t)Cout + co) * aux_cin + ci]);     // shift
                        }
                    }
        for (int n = 0; n < NT; ++n) {
            const int co = nt * NT + n;
            if (co >= Cout) break;
            float bv = b ? b[co] : 0.f;
            if (aux_mode == TC_AUX_ACC && aux_b) bv += aux_b[co];
            hb[(size_t)nt * o.NTp + n] = bv;
            if (aux_mode == TC_AUX_FILM && aux_b) {
                hf[(size_t)nt * film_rows + n] = aux_b[co];
                hf[(size_t)nt * film_rows + o.NTp + n] = aux_b[Cout + co];
            }
        }
    }
    TVC_CUDA(cudaMalloc(&o.w, img.size() * sizeof(uint16_t)));
    TVC_CUDA(cudaMemcpy(o.w, img.data(), img.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    TVC_CUDA(cudaMalloc(&o.bias, hb.size() * sizeof(float)));
    TVC_CUDA(cudaMemcpy(o.bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (!hf.empty()) {
        TVC_CUDA(cudaMalloc(&o.film_bias, hf.size() * sizeof(float)));
        TVC_CUDA(cudaMemcpy(o.film_bias, hf.data(), hf.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

constexpr int kTcMaxSmem = 200 * 1024;
constexpr int kTcTopkSmem = 227 * 1024;         // the fused top-k kernel: a deeper ring beside its merge scratch

static int g_num_sms = 148;
bool g_force_generic = false;   // tests: run the runtime-flag instantiation
bool g_force_flat = false;      // tests: never use the shared-window (halo) mode

int tc_conv_init() {
    int dev = 0, n = 0;
    TVC_CUDA(cudaGetDevice(&dev));
    TVC_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (n > 0) g_num_sms = n;
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<-1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcTopkSmem));
    TVC_CUDA(cudaFuncSetAttribute(tc_conv_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcMaxSmem));
    return 0;
}

// ---- developer timeline (see Tracer) -------------------------------------------------------------
constexpr int kTraceKernels = 8;
static uint2* g_trace_buf = nullptr;           // [kTraceKernels][4 roles][kTraceRegion]
static int g_trace_want[kTraceKernels], g_trace_n = 0, g_trace_count = -1;   // -1: not armed

int tc_trace_arm(const char* list) {
    g_trace_n = 0;
    for (const char* c = list; *c && g_trace_n < kTraceKernels;) {
        g_trace_want[g_trace_n++] = atoi(c);
        while (*c && *c != ',') ++c;
        if (*c == ',') ++c;
    }
    const size_t bytes = sizeof(uint2) * kTraceKernels * 4 * kTraceRegion;
    if (!g_trace_buf) TVC_CUDA(cudaMalloc(&g_trace_buf, bytes));
    TVC_CUDA(cudaMemset(g_trace_buf, 0, bytes));
    g_trace_count = 0;
    return 0;
}
static uint2* tc_trace_slot() {
    if (g_trace_count < 0) return nullptr;
    const int k = g_trace_count++;
    for (int i = 0; i < g_trace_n; ++i)
        if (g_trace_want[i] == k) return g_trace_buf + (size_t)i * 4 * kTraceRegion;
    return nullptr;
}
int tc_trace_dump(const char* path) {
    TVC_REQUIRE(g_trace_buf, "tc_trace: nothing armed");
    TVC_CUDA(cudaDeviceSynchronize());
    std::vector<uint2> h((size_t)kTraceKernels * 4 * kTraceRegion);
    TVC_CUDA(cudaMemcpy(h.data(), g_trace_buf, h.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
    FILE* f = fopen(path, "w");
    TVC_REQUIRE(f, "tc_trace: cannot open %s", path);
    for (int i = 0; i < g_trace_n; ++i)
        for (int r = 0; r < 4; ++r)
            for (int e = 0; e < kTraceRegion; ++e) {
                const uint2 v = h[((size_t)i * 4 + r) * kTraceRegion + e];
                if (!v.y && !v.x) {
                    if (e < kTraceRegion - 2) e = kTraceRegion - 3;      // the kernel-exit stamps sit in the last two slots
                    continue;
                }
                fprintf(f, "%d %u %u %u %u %u\n", g_trace_want[i], v.y >> 28, (v.y >> 24) & 15u, (v.y >> 8) & 0xffffu, v.y & 0xffu, v.x);
            }
    fclose(f);
    g_trace_count = -1;
    return 0;
}

// ---- TMA tensor maps -------------------------------------------------------------------------------
typedef CUresult (*TmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmEncodeFn tm_encode_fn() {
    static TmEncodeFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (TmEncodeFn)ptr;
    }();
    return fn;
}
// Chunk-major bf16 plane with `rows` rows and `nch` 8-channel chunks, viewed as [nch][ceil(rows / 8)][64 elements]
// (8 consecutive rows of a chunk are 128 contiguous bytes); box = [box_ch][groups][64].  Row groups past the end and
// chunks past `nch` arrive as zeros (the channel padding of a K-stage needs exactly that).
int tc_make_plane_map(CUtensorMap* m, const bf16* base, long long rows, int nch, int box_ch, int groups) {
    TmEncodeFn enc = tm_encode_fn();
    TVC_REQUIRE(enc, "tc_conv: cuTensorMapEncodeTiled is not available");
    const cuuint64_t dims[3] = {64, (cuuint64_t)((rows + 7) / 8), (cuuint64_t)nch};
    const cuuint64_t strides[2] = {128, (cuuint64_t)rows * 16};
    const cuuint32_t box[3] = {64, (cuuint32_t)groups, (cuuint32_t)box_ch};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TVC_REQUIRE(r == CUDA_SUCCESS, "tc_conv: cuTensorMapEncodeTiled failed (%d) rows=%lld chunks=%d", (int)r, rows, nch);
    return 0;
}

// Ranges to cut the top-k sweep into so that (row tiles x ranges) covers the SMs: the largest divisor of the channel tiles
// (<= 8) that keeps the sweeps within one wave.  1 for anything with at least half a wave of row tiles.
int tc_conv_topk_splits(const TcConvW& W, long long rows) {
    const long long row_tiles = (rows + kTileM - 1) / kTileM;
    int best = 1;
    for (int sp = 2; sp <= 8; ++sp)
        if (W.n_tiles % sp == 0 && row_tiles * sp <= g_num_sms) best = sp;
    return best;
}

int tc_conv_launch(const TcConvW& W, const TcConvArgs& a, cudaStream_t s) {
    TVC_REQUIRE(W.w && a.a_hi && a.a_lo, "tc_conv: missing weights or input");
    TVC_REQUIRE(!W.cat, "tc_conv: \"cat\" weight images belong to the fused block kernel (tc_block.cu)");
    TVC_REQUIRE(a.B > 0 && a.T > 0, "tc_conv: empty problem B=%d T=%d", a.B, a.T);
    TVC_REQUIRE(a.a_cs % 8 == 0 && a.a_cs >= W.Cin, "tc_conv: input channel stride %d (need multiple of 8 >= %d)", a.a_cs, W.Cin);
    TVC_REQUIRE(W.aux_mode == TC_AUX_NONE || (a.x_hi && a.x_lo && a.x_cs % 8 == 0 && a.x_cs >= W.aux_cin), "tc_conv: aux input missing / bad stride");
    TVC_REQUIRE(!a.y_hi || (a.y_lo && a.y_cs % 8 == 0), "tc_conv: plane output needs both planes and a stride multiple of 8");
    TVC_REQUIRE(!a.y32 || a.y32_cs % 8 == 0, "tc_conv: fp32 output capacity must be a multiple of 8 channels");
    TVC_REQUIRE(!a.res || (a.res_cs % 8 == 0 && a.res_cs >= (int)align_up(W.Cout, 8)), "tc_conv: residual capacity must be a multiple of 8 covering Cout rounded up to 8");
    TcKParams p;
    p.a_hi = a.a_hi; p.a_lo = a.a_lo; p.x_hi = a.x_hi; p.x_lo = a.x_lo; p.w = W.w;
    p.bias = W.bias; p.film_bias = W.film_bias; p.res = a.res; p.y32 = a.y32; p.y_hi = a.y_hi; p.y_lo = a.y_lo;
    p.rows = (long long)a.B * a.T; p.tile_elems = (long long)W.tile_elems;
    TVC_REQUIRE(a.a_pad >= 0 && a.y_pad >= 0 && (a.a_pad == 0 || (W.taps == 3 && a.a_pad >= a.dil)),
                "tc_conv: input padding %d must cover the dilation %d of a k = 3 conv", a.a_pad, a.dil);
    TVC_REQUIRE(a.y_pad == 0 || (a.a_pad > 0 && a.y_hi), "tc_conv: a padded plane output needs the padded mode (a_pad > 0)");
    p.a_pad = a.a_pad; p.y_pad = a.y_pad; p.Tp = a.T + 2 * a.a_pad;
    p.rows_a = (long long)a.B * p.Tp;
    p.rows_y = (long long)a.B * (a.T + 2 * a.y_pad);
    TVC_REQUIRE(p.rows_a < (1LL << 31) && p.rows_y < (1LL << 31), "tc_conv: too many rows");
    p.a_cs = a.a_cs; p.x_cs = a.x_cs; p.res_cs = a.res_cs; p.y32_cs = a.y32_cs; p.y_cs = a.y_cs;
    p.T = a.T; p.dil = a.dil; p.taps = W.taps; p.nkb = W.nkb; p.aux_nkb = W.aux_nkb; p.aux_mode = W.aux_mode;
    p.KB = W.KB; p.NT = W.NT; p.NTp = W.NTp; p.Cout = W.Cout;
    p.epi_act = a.epi_act; p.out_act = a.out_act;
    const bool topk = a.topk_cand != nullptr;
    TVC_REQUIRE(!topk || (a.topk_flag && W.taps == 1 && W.aux_mode == TC_AUX_NONE && !a.y32 && !a.y_hi && !a.res && a.a_pad == 0 &&
                          a.topk_k >= 1 && a.topk_k <= 4 && a.topk_n >= 1 && a.topk_n <= W.Cout),
                "tc_conv: the fused top-k needs a plain 1 x 1 conv without outputs, 1 <= k <= 4");
    p.topk_cand = a.topk_cand; p.topk_flag = a.topk_flag; p.topk_k = a.topk_k; p.topk_n = a.topk_n; p.topk_eps = a.topk_eps;
    p.row_major = topk ? 1 : 0;
    p.topk_splits = topk ? a.topk_splits : 1;
    p.topk_score = a.topk_score;
    TVC_REQUIRE(p.topk_splits >= 1 && W.n_tiles % p.topk_splits == 0 && (p.topk_splits == 1 || a.topk_score),
                "tc_conv: %d top-k ranges do not divide %d channel tiles (or no score buffer)", p.topk_splits, W.n_tiles);
    p.topk_nper = W.n_tiles / p.topk_splits;
    const bool dec = a.dec_f > 0;
    TVC_REQUIRE(!dec || (a.dec_r_hi && a.dec_r_lo && a.dec_a_hi && a.dec_a_lo && a.y_hi && !a.y32 && !a.res && W.aux_mode != TC_AUX_FILM &&
                         a.epi_act == TC_ACT_NONE && a.out_act == TC_ACT_NONE && a.T % a.dec_f == 0 && a.dec_pad >= 0 &&
                         (a.dec_f == 3 || a.dec_f == 5 || (a.dec_f == 4 && a.T % 4 == 0 && (a.a_pad % 4 == 0)))),
                "tc_conv: the fused down-resampler needs a plain plane output and a factor of 3, 4 or 5 dividing T");
    p.dec_f = a.dec_f; p.dec_pad = a.dec_pad; p.dec_scale = a.dec_scale;
    p.dec_r_hi = a.dec_r_hi; p.dec_r_lo = a.dec_r_lo; p.dec_a_hi = a.dec_a_hi; p.dec_a_lo = a.dec_a_lo;
    const int topk_bytes = topk ? 3 * kTileM * kTopkC * 8 : 0;
    const int smem_budget = topk ? kTcTopkSmem : kTcMaxSmem;
    static const int dbg = getenv("TVC_TC_DBG") ? atoi(getenv("TVC_TC_DBG")) : 0;
    p.dbg = dbg;
    p.trace = tc_trace_slot();
    // halo mode: k=3 convs whose utterances fill their 128-row tiles well share one row window across the taps
    const int tpu = cdiv(a.T, kTileM);
    p.halo = (W.taps == 3 && (double)a.T / ((double)tpu * kTileM) >= 0.75 && !g_force_flat) ? 1 : 0;
    if (a.a_pad > 0) p.halo = 2;         // stored replicate padding: every window is a contiguous tensor copy
    p.tiles_per_utt = tpu;
    p.R = p.halo ? kTileM + 2 * a.dil : kTileM;
    TVC_REQUIRE(a.dil >= 1 && a.dil <= 64, "tc_conv: dilation %d out of range", a.dil);
    // Stage geometry: a main / aux stage holds G row groups of 8 rows per chunk column: the window (R or 128 rows) plus
    // up to 7 leading rows, because a TMA box starts on a multiple of 8 operand rows.
    const int g_main = p.halo ? (p.R + 7 + 7) / 8 : kTileM / 8;
    const int g_aux = p.halo == 1 ? (kTileM + 7 + 7) / 8 : kTileM / 8;      // padded mode gathers the aux rows (no offset)
    p.lbo_main = (uint32_t)g_main * 128u;
    p.lbo_aux = (uint32_t)g_aux * 128u;
    const int n_rows_max = W.aux_mode == TC_AUX_FILM ? 2 * W.NTp : W.NTp;
    const uint32_t lbo_max = (W.aux_mode && p.lbo_aux > p.lbo_main) ? p.lbo_aux : p.lbo_main;
    p.a_stage_bytes = (uint32_t)align_up(2 * (W.KB / 8) * (int)lbo_max, 128);         // 2 planes x KB/8 chunk columns
    static const int tma_env = getenv("TVC_TC_TMA") ? atoi(getenv("TVC_TC_TMA")) : 1;
    memset(&p.tm_a_hi, 0, 4 * sizeof(CUtensorMap));
    p.tma_main = (tma_env && (p.halo || W.taps == 1)) ? 1 : 0;
    p.tma_aux = (tma_env && W.aux_mode != TC_AUX_NONE && p.halo != 2) ? 1 : 0;
    TVC_REQUIRE(p.halo != 2 || p.tma_main, "tc_conv: the padded mode needs tensor copies (TVC_TC_TMA=0 is set)");
    if (p.tma_main) {
        TVC_TRY(tc_make_plane_map(&p.tm_a_hi, a.a_hi, p.rows_a, a.a_cs / 8, W.KB / 8, g_main));
        TVC_TRY(tc_make_plane_map(&p.tm_a_lo, a.a_lo, p.rows_a, a.a_cs / 8, W.KB / 8, g_main));
    }
    if (p.tma_aux) {
        TVC_TRY(tc_make_plane_map(&p.tm_x_hi, a.x_hi, p.rows, a.x_cs / 8, W.KB / 8, g_aux));
        TVC_TRY(tc_make_plane_map(&p.tm_x_lo, a.x_lo, p.rows, a.x_cs / 8, W.KB / 8, g_aux));
    }
    uint32_t b_main = 4u * (uint32_t)W.KB * (uint32_t)W.NTp * (uint32_t)(p.halo ? W.taps : 1);
    uint32_t b_aux = W.aux_mode ? 4u * (uint32_t)W.KB * (uint32_t)n_rows_max : 0u;
    p.b_stage_bytes = b_main > b_aux ? b_main : b_aux;
    // Weights-resident mode: a layer with a single channel tile whose whole weight image is small keeps it in shared
    // memory for the life of the CTA (the 24- and 48-channel layers of the two highest rates: every tile used to re-fetch
    // the same 12-46 KB); the ring then holds activation windows only.
    static const int wres_env = getenv("TVC_TC_WRES") ? atoi(getenv("TVC_TC_WRES")) : 1;
    const size_t image_bytes = W.tile_elems * sizeof(bf16);
    p.w_bytes = (wres_env && W.n_tiles == 1 && image_bytes <= 64 * 1024 && !(dbg & 1)) ? (uint32_t)image_bytes : 0u;
    if (p.w_bytes) p.b_stage_bytes = 0;
    const uint32_t stage = p.a_stage_bytes + p.b_stage_bytes;
    int ring = (smem_budget - 256 - topk_bytes - (int)p.w_bytes) / (int)stage;
    ring = ring > 8 ? 8 : ring;
    TVC_REQUIRE(ring >= 1, "tc_conv: a K-stage of %u bytes does not fit shared memory", stage);
    p.ring = ring;
    p.ring_alloc = ring;
    const uint32_t cols = 2u * (uint32_t)(W.aux_mode == TC_AUX_FILM ? 3 * W.NTp : W.NTp);   // double-buffered accumulators
    uint32_t tc = 32;
    while (tc < cols) tc <<= 1;
    TVC_REQUIRE(tc <= 512, "tc_conv: %u TMEM columns needed (> 512)", cols);
    p.tmem_cols = tc;
    p.row_tiles = p.halo == 1 ? (long long)a.B * tpu : (p.rows_a + kTileM - 1) / kTileM;
    p.n_tiles = W.n_tiles;
    p.topk_smem = (uint32_t)align_up((int64_t)ring * stage + p.w_bytes + 16 * ring + 64, 16);
    const size_t smem = topk ? (size_t)p.topk_smem + topk_bytes : (size_t)ring * stage + p.w_bytes + 16 * ring + 64;
    const long long tiles = p.row_tiles * p.n_tiles;
    const long long work = topk ? p.row_tiles * p.topk_splits : tiles;   // row-major: a CTA takes whole (row tile, range) sweeps
    const long long sms = (t_sm_cap > 0 && t_sm_cap < g_num_sms) ? t_sm_cap : g_num_sms;
    const unsigned grid = (unsigned)(work < sms ? work : sms);
    // Two MMA-issuing warps (each with its own half of the ring) when a CTA has several tiles to alternate and the ring
    // is deep enough that half of it still prefetches ahead.
    // Same-box A/B (profiles/r01z_ab_issuer_modes.log): tiles of one K-stage gain 3-6 % from the second issuer, FiLM / deep-K
    // tiles lose as much, so the default (2) uses it for single-stage tiles only; 1 = wherever eligible, 0 = never.
    static const int mma2_env = getenv("TVC_TC_MMA2") ? atoi(getenv("TVC_TC_MMA2")) : 2;
    const int stages_per_tile = (p.halo ? W.nkb : W.taps * W.nkb) + (W.aux_mode ? W.aux_nkb : 0);
    p.mma2 = (mma2_env && ring >= 4 && tiles >= 2LL * grid && (mma2_env != 2 || stages_per_tile == 1)) ? 1 : 0;   // 2: single-stage tiles only
    if (p.mma2 && (ring & 1)) p.ring = ring - 1;                   // equal halves (the smem layout keeps the full ring)
    int spec = -1;
    for (int i = 0; i < kNumSpecs; ++i) {
        const EpiSpec e = epi_spec(i);
        if ((e.topk != 0) != topk || (e.dec != 0) != dec) continue;
        if ((e.film != 0) == (W.aux_mode == TC_AUX_FILM) && (e.res != 0) == (a.res != nullptr) && (e.y32 != 0) == (a.y32 != nullptr) &&
            (e.planes != 0) == (a.y_hi != nullptr) && e.epi_act == a.epi_act && (e.out_act == a.out_act || !a.y_hi)) {
            spec = i;
            break;
        }
    }
    if (g_force_generic && !topk && !dec) spec = -1;
    TVC_REQUIRE(!topk || spec == 8, "tc_conv: no fused top-k instantiation");
    TVC_REQUIRE(!dec || spec == 9, "tc_conv: no fused down-resampler instantiation");
    switch (spec) {
        case 0: TVC_LAUNCH_PDL(tc_conv_kernel<0>, grid, kThreads, smem, s, p); break;
        case 1: TVC_LAUNCH_PDL(tc_conv_kernel<1>, grid, kThreads, smem, s, p); break;
        case 2: TVC_LAUNCH_PDL(tc_conv_kernel<2>, grid, kThreads, smem, s, p); break;
        case 3: TVC_LAUNCH_PDL(tc_conv_kernel<3>, grid, kThreads, smem, s, p); break;
        case 4: TVC_LAUNCH_PDL(tc_conv_kernel<4>, grid, kThreads, smem, s, p); break;
        case 5: TVC_LAUNCH_PDL(tc_conv_kernel<5>, grid, kThreads, smem, s, p); break;
        case 6: TVC_LAUNCH_PDL(tc_conv_kernel<6>, grid, kThreads, smem, s, p); break;
        case 7: TVC_LAUNCH_PDL(tc_conv_kernel<7>, grid, kThreads, smem, s, p); break;
        case 8: TVC_LAUNCH_PDL(tc_conv_kernel<8>, grid, kThreads, smem, s, p); break;
        case 9: TVC_LAUNCH_PDL(tc_conv_kernel<9>, grid, kThreads, smem, s, p); break;
        default: TVC_LAUNCH_PDL(tc_conv_kernel<-1>, grid, kThreads, smem, s, p); break;
    }
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc

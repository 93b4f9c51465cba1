// Generic dense Conv1d for the TinyVC hot path, fp32 on CUDA cores (exact-arithmetic path).
//
// Computes, for every utterance b, output channel co and time step t
//     y[b,co,t] = epi( bias[co] + sum_{tap<K} sum_{ci<Cin} W[co,ci,tap] * pre(x[b,ci,clamp(t+(tap-(K-1)/2)*dil)]) )
// which covers every dense convolution of the reference's inference path:
//   * all 1x1 convs (encoder.py:27,30,84,87; convnext.py:44,46; decoder.py:93-94,119-124,143,171,201)
//   * the k=3 dilated `padding_mode='replicate'` convs of Downsample/Upsample (decoder.py:143-146,165-170)
//     and `downs.0` (decoder.py:206) -- replicate padding == clamping the time index.
// `pre` fuses the producer-side element-wise op of the reference (leaky_relu 0.1, or the GRN
// affine of convnext.py:34 folded to x*s[b,ci]+h[ci]); `epi` fuses the consumer side (residual
// add, FiLM x*scale+shift+res of decoder.py:97,180-181, GELU, ELU+1).
//
// Layout: activations are the reference's channels-first [B][C][T] fp32.  The (b,t) axes are
// flattened into one column axis so that short utterances (Lf = 18 frames) still fill a tile;
// a tile may straddle utterances, every column clamps its taps inside its own utterance.
// Weights are pre-packed [K][Cin][CoutP] (CoutP = Cout rounded up to 4, zero filled) so that
// a (tap,ci) row is a contiguous, float4-aligned vector over output channels.
//
// Tiling: block = BM output channels x BN columns, K-chunks of BK input channels.  Inputs are
// staged in shared memory tap-expanded ([BK][K][BN]) so the inner loop reads aligned float4
// columns whatever the dilation; weights ([K][BK][BM]) are warp-broadcast float4 reads.
// Each thread owns a TM x 4 register tile.
#include "tvc_common.cuh"

namespace tvc {

int g_conv_impl = CONV_IMPL_TC;   // Decoder.infer runs on the tcgen05 path unless ("conv_impl","fp32") is set

template <int NTY, int TM, int NTX, int BK, int KT>
__global__ void __launch_bounds__(NTY* NTX) conv1d_f32_kernel(ConvParams p) {
    constexpr int TN = 4;
    constexpr int BM = NTY * TM, BN = NTX * TN, NT = NTY * NTX;
    constexpr int CPT = (BN >= NT) ? BN / NT : 1;    // columns a loader thread owns
    constexpr int RSTEP = (BN >= NT) ? 1 : NT / BN;  // rows covered per loader pass
    static_assert((BN >= NT) ? (BN % NT == 0) : (NT % BN == 0), "loader mapping");
    static_assert(TM % 4 == 0, "TM must be a multiple of 4");

    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                  // [BK*KT][BN]
    float* ws = smem + BK * KT * BN;   // [KT][BK][BM]

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const long long n0 = (long long)blockIdx.x * BN;
    const int co0 = blockIdx.y * BM;
    const long long ncol = (long long)p.B * p.T;
    const int T = p.T;

    // ---- loader bookkeeping: fixed column set per thread ----
    const int jbase = (BN >= NT) ? tid : tid % BN;
    const int rbase = (BN >= NT) ? 0 : tid / BN;
    long long colbase[CPT];
    int colb[CPT];
    int toff[CPT][KT];
    bool colok[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const long long n = n0 + jbase + c * NT;
        colok[c] = n < ncol;
        const int b = colok[c] ? (int)(n / T) : 0;
        const int t = colok[c] ? (int)(n - (long long)b * T) : 0;
        colb[c] = b;
        colbase[c] = (long long)b * p.x_bs;
#pragma unroll
        for (int tap = 0; tap < KT; ++tap) {
            int tt = t + (tap - (KT - 1) / 2) * p.dil;
            tt = tt < 0 ? 0 : (tt > T - 1 ? T - 1 : tt);
            toff[c][tap] = tt;
        }
    }

    float acc[TM][TN];
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;

    const int pre = p.pre;
    // Software pipeline: chunk c+1 travels global -> registers while chunk c is multiplied out of shared memory.
    constexpr int XROWS = (BK * KT + RSTEP - 1) / RSTEP;          // staged input rows per loader thread
    constexpr int WPT = (KT * BK * (BM / 4) + NT - 1) / NT;       // staged weight float4s per thread
    // Loads are straight-line (indices clamped to legal addresses, masking and the producer-side op applied when the chunk
    // is stored to shared memory): a value consumed inside its own guarded region would serialise the chunk's loads, one
    // global-memory round trip per element (measured: 14 us per 32-row chunk with PRE_AFFINE on a streaming tick).
    constexpr int SR = (KT == 1) ? XROWS : 1;                     // GRN affine operands exist for 1x1 convs only
    float xr[XROWS][CPT];
    float sr[SR][CPT], hr[SR];
    float4 wr[WPT];
    const int cin_last = p.Cin - 1;
    auto load_chunk = [&](int ci0) {
#pragma unroll
        for (int r = 0; r < XROWS; ++r) {
            const int row = min(rbase + r * RSTEP, BK * KT - 1);
            const int k = row / KT, tap = row - k * KT;
            const int ci = min(ci0 + k, cin_last);
#pragma unroll
            for (int c = 0; c < CPT; ++c) xr[r][c] = __ldg(p.x + colbase[c] + (long long)ci * T + toff[c][tap]);
        }
        if (KT == 1 && pre == PRE_AFFINE) {
#pragma unroll
            for (int r = 0; r < SR; ++r) {
                const int ci = min(ci0 + min(rbase + r * RSTEP, BK - 1), cin_last);
                hr[r] = __ldg(p.pre_shift + ci);
#pragma unroll
                for (int c = 0; c < CPT; ++c) sr[r][c] = __ldg(p.pre_scale + (long long)colb[c] * p.Cin + ci);
            }
        }
#pragma unroll
        for (int q = 0; q < WPT; ++q) {
            const int e = min(tid + q * NT, KT * BK * (BM / 4) - 1);
            const int m4 = e % (BM / 4);
            const int rk = e / (BM / 4);          // tap*BK + k
            const int tap = rk / BK, k = rk - tap * BK;
            const int ci = ci0 + k, co = co0 + m4 * 4;
            const bool ok = ci < p.Cin && co < p.CoutP;
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + ((long long)tap * p.Cin + min(ci, cin_last)) * p.CoutP + (ok ? co : 0)));
            wr[q] = ok ? v : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto store_chunk = [&](int ci0) {
#pragma unroll
        for (int r = 0; r < XROWS; ++r) {
            const int row = rbase + r * RSTEP;
            if (row < BK * KT) {
                const int ci = ci0 + row / KT;
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    float v = xr[r][c];
                    if (pre == PRE_LRELU) v = leaky01(v);
                    else if (KT == 1 && pre == PRE_AFFINE) v = fmaf(v, sr[r < SR ? r : 0][c], hr[r < SR ? r : 0]);
                    xs[row * BN + jbase + c * NT] = (colok[c] && ci < p.Cin) ? v : 0.f;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < WPT; ++q) {
            const int e = tid + q * NT;
            if (e < KT * BK * (BM / 4)) {
                const int m4 = e % (BM / 4);
                const int rk = e / (BM / 4);
                *reinterpret_cast<float4*>(ws + rk * BM + m4 * 4) = wr[q];
            }
        }
    };
    load_chunk(0);
    for (int ci0 = 0; ci0 < p.Cin; ci0 += BK) {
        store_chunk(ci0);
        __syncthreads();
        if (ci0 + BK < p.Cin) load_chunk(ci0 + BK);
        // ---- FMA (per output: acc = fma(w, x, acc) with ci ascending, taps inner -- the order parity tests pin) ----
#pragma unroll
        for (int k = 0; k < BK; ++k) {
#pragma unroll
            for (int tap = 0; tap < KT; ++tap) {
                const float4 xv = *reinterpret_cast<const float4*>(xs + (k * KT + tap) * BN + tx * TN);
                const float* wrow = ws + (tap * BK + k) * BM + ty * TM;
#pragma unroll
                for (int m4 = 0; m4 < TM / 4; ++m4) {
                    const float4 wv = *reinterpret_cast<const float4*>(wrow + m4 * 4);
                    const float wvv[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc[m4 * 4 + q][0] = fmaf(wvv[q], xv.x, acc[m4 * 4 + q][0]);
                        acc[m4 * 4 + q][1] = fmaf(wvv[q], xv.y, acc[m4 * 4 + q][1]);
                        acc[m4 * 4 + q][2] = fmaf(wvv[q], xv.z, acc[m4 * 4 + q][2]);
                        acc[m4 * 4 + q][3] = fmaf(wvv[q], xv.w, acc[m4 * 4 + q][3]);
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue ----
    const long long nc0 = n0 + tx * TN;
    if (nc0 >= ncol) return;
    const int epi = p.epi;
    const bool vec = (T % 4 == 0);   // then the 4 columns share an utterance and are 16B aligned
    if (vec) {
        const int b = (int)(nc0 / T);
        const int t = (int)(nc0 - (long long)b * T);
#pragma unroll
        for (int m = 0; m < TM; ++m) {
            const int co = co0 + ty * TM + m;
            if (co >= p.Cout) break;
            const float bv = p.bias ? __ldg(p.bias + co) : 0.f;
            float o[4] = {acc[m][0] + bv, acc[m][1] + bv, acc[m][2] + bv, acc[m][3] + bv};
            if (epi == EPI_RES) {
                const float4 r = *reinterpret_cast<const float4*>(p.res + (long long)b * p.res_bs + (long long)co * T + t);
                o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
            } else if (epi == EPI_FILM_RES) {
                const float* fb = p.film + (long long)b * p.film_bs + t;
                const float4 sc = *reinterpret_cast<const float4*>(fb + (long long)co * T);
                const float4 sh = *reinterpret_cast<const float4*>(fb + (long long)(co + p.Cout) * T);
                const float4 r = *reinterpret_cast<const float4*>(p.res + (long long)b * p.res_bs + (long long)co * T + t);
                o[0] = __fadd_rn(__fadd_rn(__fmul_rn(o[0], sc.x), sh.x), r.x);
                o[1] = __fadd_rn(__fadd_rn(__fmul_rn(o[1], sc.y), sh.y), r.y);
                o[2] = __fadd_rn(__fadd_rn(__fmul_rn(o[2], sc.z), sh.z), r.z);
                o[3] = __fadd_rn(__fadd_rn(__fmul_rn(o[3], sc.w), sh.w), r.w);
            } else if (epi == EPI_GELU) {
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = gelu_erf(o[q]);
            } else if (epi == EPI_ELU1) {
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = elu_plus1(o[q]);
            }
            *reinterpret_cast<float4*>(p.y + (long long)b * p.y_bs + (long long)co * T + t) = make_float4(o[0], o[1], o[2], o[3]);
        }
    } else {
#pragma unroll
        for (int n = 0; n < TN; ++n) {
            const long long nc = nc0 + n;
            if (nc >= ncol) break;
            const int b = (int)(nc / T);
            const int t = (int)(nc - (long long)b * T);
#pragma unroll
            for (int m = 0; m < TM; ++m) {
                const int co = co0 + ty * TM + m;
                if (co >= p.Cout) break;
                float o = acc[m][n] + (p.bias ? __ldg(p.bias + co) : 0.f);
                if (epi == EPI_RES) {
                    o += p.res[(long long)b * p.res_bs + (long long)co * T + t];
                } else if (epi == EPI_FILM_RES) {
                    const float* fb = p.film + (long long)b * p.film_bs + t;
                    const float sc = fb[(long long)co * T], sh = fb[(long long)(co + p.Cout) * T];
                    o = __fadd_rn(__fadd_rn(__fmul_rn(o, sc), sh), p.res[(long long)b * p.res_bs + (long long)co * T + t]);
                } else if (epi == EPI_GELU) {
                    o = gelu_erf(o);
                } else if (epi == EPI_ELU1) {
                    o = elu_plus1(o);
                }
                p.y[(long long)b * p.y_bs + (long long)co * T + t] = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
template <int NTY, int TM, int NTX, int BK, int KT>
struct ConvCfg {
    static constexpr int BM = NTY * TM, BN = NTX * 4, NT = NTY * NTX;
    static constexpr size_t smem = sizeof(float) * (size_t)(BK * KT * BN + KT * BK * BM);
    static int init() {
        TVC_CUDA(cudaFuncSetAttribute(conv1d_f32_kernel<NTY, TM, NTX, BK, KT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
    }
    static int launch(const ConvParams& p, cudaStream_t s) {
        const long long ncol = (long long)p.B * p.T;
        dim3 grid((unsigned)((ncol + BN - 1) / BN), (unsigned)((p.Cout + BM - 1) / BM), 1);
        conv1d_f32_kernel<NTY, TM, NTX, BK, KT><<<grid, NT, smem, s>>>(p);
        TVC_LAUNCH_CHECK();
        return 0;
    }
};

// Tile families: F24/F48/F96 for the FilterNet's 24*k channel counts, G64 general.
template <int KT> using F24 = ConvCfg<2, 12, 64, 8, KT>;    // 24 x 256, 128 threads
template <int KT> using F48 = ConvCfg<4, 12, 64, 8, KT>;    // 48 x 256, 256 threads
template <int KT> using F96 = ConvCfg<8, 12, 32, 8, KT>;    // 96 x 128, 256 threads
template <int KT> using G64 = ConvCfg<8, 8, 32, 8, KT>;     // 64 x 128, 256 threads
template <int KT> using G16 = ConvCfg<2, 8, 64, 8, KT>;     // 16 x 256, 128 threads (tiny Cout)
using G128 = ConvCfg<8, 16, 32, 16, 1>;                      // 128 x 128, 256 threads, BK 16: the big 1x1 products (encoder, STFT, kNN)
// 32 x 64, 64 threads, BK 32: 1x1 products over few columns (a streaming tick's PitchEstimator: 3 584 frames).  Four times
// the CTAs of G64 and a quarter of the K-chunks, i.e. of the load -> stage -> multiply round trips that bound a CTA there;
// the per-output FMA order (ci ascending) and therefore the result bits are those of every other configuration.
using S32 = ConvCfg<4, 8, 16, 32, 1>;

int conv1d_init() {
    TVC_TRY(F24<1>::init()); TVC_TRY(F24<3>::init());
    TVC_TRY(F48<1>::init()); TVC_TRY(F48<3>::init());
    TVC_TRY(F96<1>::init()); TVC_TRY(F96<3>::init());
    TVC_TRY(G64<1>::init()); TVC_TRY(G64<3>::init());
    TVC_TRY(G16<1>::init()); TVC_TRY(G16<3>::init());
    TVC_TRY(G128::init());
    TVC_TRY(S32::init());
    return 0;
}

template <int KT>
static int dispatch_cfg(const ConvParams& p, cudaStream_t s) {
    const int co = p.Cout;
    if (co <= 16) return G16<KT>::launch(p, s);
    if (co <= 24) return F24<KT>::launch(p, s);
    if (co <= 48) return F48<KT>::launch(p, s);
    if (co % 96 == 0) return F96<KT>::launch(p, s);
    return G64<KT>::launch(p, s);
}

int conv1d_launch(const ConvParams& p, cudaStream_t stream) {
    TVC_REQUIRE(p.B > 0 && p.T > 0 && p.Cin > 0 && p.Cout > 0, "conv1d: empty problem B=%d T=%d Cin=%d Cout=%d", p.B, p.T, p.Cin, p.Cout);
    TVC_REQUIRE(p.CoutP % 4 == 0 && p.CoutP >= p.Cout, "conv1d: CoutP=%d must be a multiple of 4 >= Cout=%d", p.CoutP, p.Cout);
    TVC_REQUIRE(p.epi != EPI_FILM_RES || (p.film && p.res), "conv1d: FiLM epilogue needs film and res");
    TVC_REQUIRE(p.epi != EPI_RES || p.res, "conv1d: residual epilogue needs res");
    TVC_REQUIRE(p.pre != PRE_AFFINE || (p.pre_scale && p.pre_shift && p.K == 1), "conv1d: affine prologue needs scale/shift and a 1x1 conv");
    // large 1x1 products: 128 x 128 tiles when they still give every SM at least ~2 CTAs
    if (p.K == 1 && p.Cout >= 256) {
        const long long tiles = (((long long)p.B * p.T + 127) / 128) * ((p.Cout + 127) / 128);
        if (tiles >= 296) return G128::launch(p, stream);
    }
    if (p.K == 1 && p.Cout > 48 && p.Cout % 32 == 0) {
        const long long g64 = (((long long)p.B * p.T + 127) / 128) * ((p.Cout + 63) / 64);
        if (g64 < 2 * 148) return S32::launch(p, stream);
    }
    if (p.K == 1) return dispatch_cfg<1>(p, stream);
    if (p.K == 3) return dispatch_cfg<3>(p, stream);
    set_error("conv1d: unsupported kernel size %d", p.K);
    return 2;
}

}  // namespace tvc

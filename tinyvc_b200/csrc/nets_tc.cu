// Decoder.infer on the tensor-core path: every dense conv of SourceNet / dsp / FilterNet
// (module/tinyvc/decoder.py:102-257) runs on tc_conv_kernel (tcgen05, split bf16 x3, fp32 TMEM
// accumulate); activations stay channels-last, chunk-major (tc_conv.cuh) between layers (fp32 where a
// residual or a resampler needs the exact value, split bf16 planes where the consumer is a conv).
#include "nets_tc.cuh"
#include "tc_block.cuh"

#include <cmath>
#include <string>
#include <vector>

namespace tvc {

namespace {

const int kUpCh[5] = {384, 192, 96, 48, 24};      // decoder.py:195
const int kUpOut[5] = {192, 96, 48, 24, 24};
const int kUpFac[5] = {2, 3, 4, 4, 5};            // decoder.py:196
const int kDownIn[4] = {24, 48, 96, 192};
const int kDownOut[4] = {48, 96, 192, 384};
const int kDownFac[4] = {5, 4, 4, 3};
// output channels per CTA (tc_conv N tile), chosen so the low-rate levels still spread over the SMs
const int kDownNT12[4] = {24, 48, 32, 32};
const int kDownNT3[4] = {48, 48, 64, 48};
const int kUpNT[5] = {48, 64, 48, 48, 24};
const int kUpNT5[5] = {32, 32, 48, 24, 24};

constexpr int kHeadsCout = 983;    // kernel 961 | 7 zero rows | amps 15  (amps start on a 16-byte chunk)
constexpr int kHeadsCs = 984;
constexpr int kAmpsOff = 968;
constexpr int kBinsCs = 968;
constexpr int kFrameInK = kContent + 2;
constexpr int kFrameInCs = 776;

struct Pl {
    bf16 *hi = nullptr, *lo = nullptr;
    int cs = 0;
};
Pl planes(Arena& A, long long rows, int cs) {
    Pl p;
    p.hi = (bf16*)A.bytes((size_t)rows * cs * sizeof(bf16));
    p.lo = (bf16*)A.bytes((size_t)rows * cs * sizeof(bf16));
    p.cs = cs;
    return p;
}

struct HostW {
    std::vector<float> flat;
    const ParamTable* table = nullptr;
    const float* get(const std::string& name, int64_t* numel = nullptr) const {
        const ParamSpec* s = table->find(name);
        if (!s) return nullptr;
        if (numel) *numel = s->numel;
        return flat.data() + s->offset;
    }
};

int pack_named(const HostW& H, const std::string& prefix, int NT, TcConvW& out, const std::string& aux_prefix = "",
               int aux_mode = TC_AUX_NONE, bool cat = false) {
    const ParamSpec* w = H.table->find(prefix + ".weight");
    TVC_REQUIRE(w && H.get(prefix + ".bias"), "tc weights: no conv named %s", prefix.c_str());
    const int Cout = w->d0, Cin = w->d1, K = w->d2;
    if (aux_mode == TC_AUX_NONE) return tc_pack_conv(H.get(prefix + ".weight"), H.get(prefix + ".bias"), Cout, Cin, K, nullptr, nullptr, 0, 0, NT, out, cat);
    if (aux_mode == TC_AUX_ACC) {
        const ParamSpec* aw = H.table->find(aux_prefix + ".weight");
        TVC_REQUIRE(aw && aw->d0 == Cout && aw->d2 == 1, "tc weights: bad residual conv %s", aux_prefix.c_str());
        return tc_pack_conv(H.get(prefix + ".weight"), H.get(prefix + ".bias"), Cout, Cin, K, H.get(aux_prefix + ".weight"),
                            H.get(aux_prefix + ".bias"), aw->d1, TC_AUX_ACC, NT, out, cat);
    }
    // FiLM: aux_prefix.to_scale / aux_prefix.to_shift, both Conv1d(C, C, 1)  (decoder.py:88-97)
    const ParamSpec* sw = H.table->find(aux_prefix + ".to_scale.weight");
    const ParamSpec* hw = H.table->find(aux_prefix + ".to_shift.weight");
    TVC_REQUIRE(sw && hw && sw->d0 == Cout && hw->d0 == Cout && sw->d1 == hw->d1, "tc weights: bad FiLM %s", aux_prefix.c_str());
    const int ac = sw->d1;
    std::vector<float> fw((size_t)2 * Cout * ac), fb((size_t)2 * Cout);
    memcpy(fw.data(), H.get(aux_prefix + ".to_scale.weight"), sizeof(float) * Cout * ac);
    memcpy(fw.data() + (size_t)Cout * ac, H.get(aux_prefix + ".to_shift.weight"), sizeof(float) * Cout * ac);
    memcpy(fb.data(), H.get(aux_prefix + ".to_scale.bias"), sizeof(float) * Cout);
    memcpy(fb.data() + Cout, H.get(aux_prefix + ".to_shift.bias"), sizeof(float) * Cout);
    return tc_pack_conv(H.get(prefix + ".weight"), H.get(prefix + ".bias"), Cout, Cin, K, fw.data(), fb.data(), ac, TC_AUX_FILM, NT, out, cat);
}

}  // namespace

DecoderTC::~DecoderTC() {
    frame_in.free_all(); heads.free_all(); dft_cos.free_all(); dft_sin.free_all(); down0.free_all();
    dft_cos_w.free_all(); dft_sin_w.free_all();
    if (fork) cudaEventDestroy(fork);
    if (join) cudaEventDestroy(join);
    if (scan_ev) cudaEventDestroy(scan_ev);
    if (up0_ev) cudaEventDestroy(up0_ev);
    if (side) cudaStreamDestroy(side);
    for (auto& m : mid) { m.c2.free_all(); m.c3.free_all(); }
    for (auto& d : down) { d.c1.free_all(); d.c2.free_all(); d.c3.free_all(); }
    for (auto& u : up) { u.c1.free_all(); u.c2.free_all(); u.c3.free_all(); u.c4.free_all(); u.c5.free_all(); }
    for (auto& u : up_w) { u.c1.free_all(); u.c2.free_all(); u.c3.free_all(); u.c4.free_all(); }
    up4_cat.c1.free_all(); up4_cat.c2.free_all(); up4_cat.c3.free_all(); up4_cat.c4.free_all(); up4_cat.c5.free_all();
    if (w7_buf) cudaFree(w7_buf);
    if (rng_state) cudaFree(rng_state);
}

int DecoderTC::init(const WeightStore& store) {
    HostW H;
    H.table = &store.table;
    H.flat.resize((size_t)store.table.total);
    TVC_CUDA(cudaMemcpy(H.flat.data(), store.flat, sizeof(float) * H.flat.size(), cudaMemcpyDeviceToHost));
    const std::string sn = "source_net", fn = "filter_net";

    // ---- frame_in: content_in (+ energy_in + f0_in) of both nets as one product over [content | e_fr | log f0]
    {
        const int Cout = 128 + 384, K = kFrameInK;
        std::vector<float> w((size_t)Cout * K, 0.f), b((size_t)Cout, 0.f);
        const float *sw = H.get(sn + ".content_in.weight"), *sb = H.get(sn + ".content_in.bias");
        const float *ew = H.get(sn + ".energy_in.weight"), *eb = H.get(sn + ".energy_in.bias");
        const float *fw = H.get(sn + ".f0_in.weight"), *fb = H.get(sn + ".f0_in.bias");
        const float *cw = H.get(fn + ".content_in.weight"), *cb = H.get(fn + ".content_in.bias");
        const float *gw = H.get(fn + ".f0_in.weight"), *gb = H.get(fn + ".f0_in.bias");
        TVC_REQUIRE(sw && sb && ew && eb && fw && fb && cw && cb && gw && gb, "tc weights: missing frame-rate input convs");
        for (int co = 0; co < 128; ++co) {
            memcpy(&w[(size_t)co * K], sw + (size_t)co * kContent, sizeof(float) * kContent);
            w[(size_t)co * K + kContent] = ew[co];
            w[(size_t)co * K + kContent + 1] = fw[co];
            b[co] = (sb[co] + eb[co]) + fb[co];            // decoder.py:128 adds in this order
        }
        for (int co = 0; co < 384; ++co) {
            memcpy(&w[(size_t)(128 + co) * K], cw + (size_t)co * kContent, sizeof(float) * kContent);
            w[(size_t)(128 + co) * K + kContent + 1] = gw[co];
            b[128 + co] = cb[co] + gb[co];                 // decoder.py:223
        }
        TVC_TRY(tc_pack_conv(w.data(), b.data(), Cout, K, 1, nullptr, nullptr, 0, 0, 64, frame_in));
    }
    // ---- SourceNet ConvNeXt blocks
    {
        std::vector<float> w7((size_t)3 * 7 * 128);
        for (int i = 0; i < 3; ++i) {
            const std::string p = sn + ".mid_layers." + std::to_string(i);
            const float* dw = H.get(p + ".c1.weight");     // [128][1][7]
            TVC_REQUIRE(dw, "tc weights: missing %s.c1", p.c_str());
            for (int c = 0; c < 128; ++c)
                for (int j = 0; j < 7; ++j) w7[((size_t)i * 7 + j) * 128 + c] = dw[c * 7 + j];
        }
        TVC_CUDA(cudaMalloc(&w7_buf, sizeof(float) * w7.size()));
        TVC_CUDA(cudaMemcpy(w7_buf, w7.data(), sizeof(float) * w7.size(), cudaMemcpyHostToDevice));
        for (int i = 0; i < 3; ++i) {
            const std::string p = sn + ".mid_layers." + std::to_string(i);
            mid[i].w7 = w7_buf + (size_t)i * 7 * 128;
            mid[i].wb = store.raw(p + ".c1.bias");
            mid[i].ln_g = store.raw(p + ".norm.gamma"); mid[i].ln_b = store.raw(p + ".norm.beta");
            mid[i].grn_g = store.raw(p + ".grn.gamma"); mid[i].grn_b = store.raw(p + ".grn.beta");
            TVC_REQUIRE(mid[i].wb && mid[i].ln_g && mid[i].ln_b && mid[i].grn_g && mid[i].grn_b, "tc weights: incomplete %s", p.c_str());
            TVC_TRY(pack_named(H, p + ".c2", 64, mid[i].c2));
            TVC_TRY(pack_named(H, p + ".c3", 32, mid[i].c3));
        }
    }
    // ---- heads: to_kernel rows [0,961), to_amps rows [968,983)
    {
        std::vector<float> w((size_t)kHeadsCout * 128, 0.f), b((size_t)kHeadsCout, 0.f);
        const float *kw = H.get(sn + ".to_kernel.weight"), *kb = H.get(sn + ".to_kernel.bias");
        const float *aw = H.get(sn + ".to_amps.weight"), *ab = H.get(sn + ".to_amps.bias");
        TVC_REQUIRE(kw && kb && aw && ab, "tc weights: missing SourceNet heads");
        memcpy(w.data(), kw, sizeof(float) * kBins * 128);
        memcpy(b.data(), kb, sizeof(float) * kBins);
        memcpy(w.data() + (size_t)kAmpsOff * 128, aw, sizeof(float) * kOsc * 128);
        memcpy(b.data() + kAmpsOff, ab, sizeof(float) * kOsc);
        TVC_TRY(tc_pack_conv(w.data(), b.data(), kHeadsCout, 128, 1, nullptr, nullptr, 0, 0, 64, heads));
    }
    // ---- inverse real-DFT bases (fp64 on the host, rounded once):  C[p] = sum_f w_f cos(2 pi f p / N) / N * Re Y[f]
    {
        std::vector<float> wc((size_t)kBins * kBins), ws((size_t)kBins * kBins);
        for (int p = 0; p < kBins; ++p)
            for (int f = 0; f < kBins; ++f) {
                const double wf = (f == 0 || f == kNfft / 2) ? 1.0 : 2.0;
                const int r = (int)(((long long)f * p) % kNfft);
                const double ang = 2.0 * M_PI * (double)r / (double)kNfft;
                wc[(size_t)p * kBins + f] = (float)(wf * std::cos(ang) / (double)kNfft);
                ws[(size_t)p * kBins + f] = (float)(wf * std::sin(ang) / (double)kNfft);
            }
        TVC_TRY(tc_pack_conv(wc.data(), nullptr, kBins, kBins, 1, nullptr, nullptr, 0, 0, 64, dft_cos));
        TVC_TRY(tc_pack_conv(ws.data(), nullptr, kBins, kBins, 1, nullptr, nullptr, 0, 0, 64, dft_sin));
        TVC_TRY(tc_pack_conv(wc.data(), nullptr, kBins, kBins, 1, nullptr, nullptr, 0, 0, 128, dft_cos_w));
        TVC_TRY(tc_pack_conv(ws.data(), nullptr, kBins, kBins, 1, nullptr, nullptr, 0, 0, 128, dft_sin_w));
        TVC_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        TVC_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        TVC_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        TVC_CUDA(cudaEventCreateWithFlags(&scan_ev, cudaEventDisableTiming));
        TVC_CUDA(cudaEventCreateWithFlags(&up0_ev, cudaEventDisableTiming));
    }
    // ---- FilterNet
    TVC_TRY(pack_named(H, fn + ".downs.0", 24, down0));
    for (int i = 0; i < 4; ++i) {
        const std::string p = fn + ".downs." + std::to_string(i + 1);
        TVC_TRY(pack_named(H, p + ".c1", kDownNT12[i], down[i].c1));
        TVC_TRY(pack_named(H, p + ".c2", kDownNT12[i], down[i].c2));
        TVC_TRY(pack_named(H, p + ".c3", kDownNT3[i], down[i].c3, p + ".down_res", TC_AUX_ACC));
    }
    for (int i = 0; i < 5; ++i) {
        const std::string p = fn + ".ups." + std::to_string(i);
        TVC_TRY(pack_named(H, p + ".c1", kUpNT[i], up[i].c1));
        TVC_TRY(pack_named(H, p + ".c2", kUpNT[i], up[i].c2, p + ".film1", TC_AUX_FILM));
        TVC_TRY(pack_named(H, p + ".c3", kUpNT[i], up[i].c3));
        TVC_TRY(pack_named(H, p + ".c4", kUpNT[i], up[i].c4, p + ".film2", TC_AUX_FILM));
        TVC_TRY(pack_named(H, p + ".c5", kUpNT5[i], up[i].c5));
        if (i < 3) {        // wide variants (up_w): same arithmetic per output channel, other tiling
            TVC_TRY(pack_named(H, p + ".c1", 96, up_w[i].c1));
            TVC_TRY(pack_named(H, p + ".c3", 96, up_w[i].c3));
            if (i < 2) {
                TVC_TRY(pack_named(H, p + ".c2", 80, up_w[i].c2, p + ".film1", TC_AUX_FILM));
                TVC_TRY(pack_named(H, p + ".c4", 80, up_w[i].c4, p + ".film2", TC_AUX_FILM));
            }
        }
        if (i == 4) {       // the fused block kernel's own images of the same convs ("cat" layout, tc_conv.cuh)
            TVC_TRY(pack_named(H, p + ".c1", kUpNT[i], up4_cat.c1, "", TC_AUX_NONE, true));
            TVC_TRY(pack_named(H, p + ".c2", kUpNT[i], up4_cat.c2, p + ".film1", TC_AUX_FILM, true));
            TVC_TRY(pack_named(H, p + ".c3", kUpNT[i], up4_cat.c3, "", TC_AUX_NONE, true));
            TVC_TRY(pack_named(H, p + ".c4", kUpNT[i], up4_cat.c4, p + ".film2", TC_AUX_FILM, true));
            TVC_TRY(pack_named(H, p + ".c5", kUpNT5[i], up4_cat.c5, "", TC_AUX_NONE, true));
        }
    }
    out_w = store.raw(fn + ".output_layer.weight");
    out_b = store.raw(fn + ".output_layer.bias");
    TVC_REQUIRE(out_w && out_b, "tc weights: missing output layer");
    TVC_CUDA(cudaMalloc(&rng_state, 2 * sizeof(unsigned long long)));
    TVC_CUDA(cudaMemset(rng_state, 0, 2 * sizeof(unsigned long long)));
    ready = true;
    return 0;
}

#define RUN(expr)                     \
    do {                              \
        if (!A.dry) {                 \
            ProfScope ps__(#expr, s); \
            TVC_TRY(expr);            \
        }                             \
    } while (0)
#define ARENA_OK() TVC_REQUIRE(!A.overflow, "workspace too small: need at least %zu bytes, got %zu", A.peak, A.cap)

namespace {
// tvc_set_option("fused_up", "0"): run the x5 resampler, the 24-channel Upsample block and the output layer as separate
// launches instead of the fused block kernel (tc_block.cu).  The fused kernel evaluates x_hi * [w_hi | w_lo] as one MMA, so the
// two agree to the rounding of one fp32 addition per accumulator, not bit for bit (tests/test_gpu_fused_block.py).
bool g_fused_up = true;
// tvc_set_option("pad_max_t", "N"): levels whose utterances have at most N rows keep the replicate padding of their
// k = 3 convs STORED in the activation planes (tc_conv.cuh, padded mode), so that every operand window is one contiguous
// tensor copy.  Short utterances (config 2: 36 / 108 / 432 rows at the three lowest rates) otherwise gather their clamped
// windows row by row with cp.async, tap by tap.  0 = never.  Separate limits for the Downsample blocks (pads 1 / 2 / 4 rows:
// the row count hardly grows) and the Upsample blocks (pads up to 27 rows: more row tiles re-stream the weights).
bool g_wide_tiles = true;     // tvc_set_option("wide_tiles", "0"): always the narrow channel tiles of ups.0 / ups.1
bool g_side_branch = true;    // tvc_set_option("side_branch", "0"): no forked branch inside a decoder step (short batches otherwise run the two inverse-DFT
                              // products side by side and give the oscillator's f0-only scan and ups.0's resampler + c1 a head start)
bool g_prune_levels = true;   // tvc_set_option("prune_levels", "0"): output pruning stops at the fused block (the Upsample levels below it run in full)
bool g_fuse_down = true;      // tvc_set_option("fuse_down", "0"): separate interp_cl launches in front of the Downsample blocks
int g_pad_max_t = 0, g_pad_down_max_t = 512;     // same-box A/B (profiles/r02f_pad_ab.log): down 1.010 -> 1.004 ms, up 1.010 -> 1.032 ms
struct ConvCall {
    TcConvArgs a;
    ConvCall(const Pl& in, int B, int T, int dil = 1) {
        a.a_hi = in.hi; a.a_lo = in.lo; a.a_cs = in.cs; a.B = B; a.T = T; a.dil = dil;
    }
    ConvCall& aux(const Pl& x) { a.x_hi = x.hi; a.x_lo = x.lo; a.x_cs = x.cs; return *this; }
    ConvCall& res(const float* r, int cs) { a.res = r; a.res_cs = cs; return *this; }
    ConvCall& f32(float* y, int cs) { a.y32 = y; a.y32_cs = cs; return *this; }
    ConvCall& out(const Pl& y, int act) { a.y_hi = y.hi; a.y_lo = y.lo; a.y_cs = y.cs; a.out_act = act; return *this; }
    ConvCall& epi(int act) { a.epi_act = act; return *this; }
    ConvCall& pad(int in, int out) { a.a_pad = in; a.y_pad = out; return *this; }
};
// Narrow or wide image of the same conv for `rows` output rows: estimated time = waves x bytes a tile streams per K-stage (its
// activation window, 128 + 2 dil rows, + its weight tile; the FiLM pair counts its second accumulator's weights).
const TcConvW& pick_tiling(const TcConvW& narrow, const TcConvW& wide, long long rows, int dil) {
    if (!g_wide_tiles || !wide.w) return narrow;
    // many waves: the tile quantisation the wide tiles fix no longer matters, and their FiLM stages leave a two-deep ring
    // (measured: 64 x 500 frames 16.51 -> 16.61 ms with wide tiles, a 128-stream tick 2.77 -> 2.65 ms)
    if (((rows + 127) / 128) * narrow.n_tiles > 6 * 148) return narrow;
    auto cost = [&](const TcConvW& W) {
        const long long tiles = ((rows + 127) / 128) * W.n_tiles;
        const long long waves = (tiles + 147) / 148;
        const double a = 128.0 + 2.0 * dil, b = (W.aux_mode == TC_AUX_FILM ? 1.7 : 1.0) * W.NTp;
        return (double)waves * (a + b);
    };
    return cost(wide) < cost(narrow) ? wide : narrow;
}
int tc_conv_k(const char* name, const TcConvW& W, const ConvCall& c, cudaStream_t s) {
    ProfScope ps(name, s);
    return tc_conv_launch(W, c.a, s);
}
}  // namespace
#define CONV(name, W, call)                                   \
    do {                                                      \
        if (!A.dry) TVC_TRY(tc_conv_k(name, W, call, s));     \
    } while (0)

// Rows [wa[i], wb[i]) of every utterance that Upsample level i (0 .. 4, 4 = the fused full-rate block) has to produce when only
// the output samples [out_t0, out_t1) are needed (out_t1 < 0: everything).  Level 4 is always "full" here (the block prunes by
// windows itself); the levels below get the rows the level above reads, widened by 40 rows, or everything (see DecoderTC::infer).
// Host arithmetic only: exported as tvc_decoder_plan_windows so that the CPU tests can check it against a brute-force
// dependency trace.
void decoder_plan_windows(int Lf, int out_t0, int out_t1, bool prune_enabled, int wa[5], int wb[5]) {
    const int L = Lf * kFrame;
    int Tlev[5];
    int t = Lf;
    for (int i = 0; i < 5; ++i) { t *= kUpFac[i]; Tlev[i] = t; wa[i] = 0; wb[i] = t; }
    const bool prune = prune_enabled && out_t1 >= 0 && (out_t0 > 0 || out_t1 < L);
    if (!prune) return;
    // rows of x4 (= level 3's output) the walked windows of the block resample: window k covers block rows
    // k * 426 - 44 .. k * 426 - 43 + 512 (tc_block.cu), clamped to the utterance
    const int k_lo = out_t0 / 426, k_hi = (out_t1 - 1) / 426;
    const int t_min = std::max(0, k_lo * 426 - 44), t_max = std::min(L - 1, k_hi * 426 - 43 + 512);
    int na = std::max(0, t_min / 5 - 1), nb = std::min(Tlev[3], t_max / 5 + 2);        // needed rows of level 3
    for (int i = 3; i >= 0; --i) {
        const int ra = std::max(0, na - 40), rb = std::min(Tlev[i], nb + 40);
        if ((rb - ra) * 100 > Tlev[i] * 85 || rb - ra < 384) break;                    // this level and all below: full
        wa[i] = ra; wb[i] = rb;
        const int f = kUpFac[i];
        na = std::max(0, ra / f - 1);
        nb = std::min(i > 0 ? Tlev[i - 1] : Lf, (rb - 1) / f + 2);
    }
}

int DecoderTC::infer(Arena& A, cudaStream_t s, const float* content, const float* f0, const float* energy,
                     const float* rand01, float* out, int B, int Lf, int out_t0, int out_t1) const {
    const int L = Lf * kFrame;
    const long long rowsF = (long long)B * Lf, rowsL = (long long)B * L;
    const size_t m0 = A.mark();

    // ---- frame-rate inputs -> [SourceNet x (128) | FilterNet x0 (384)], channels-last fp32 [rowsF][512]
    float* e_fr = A.f32(rowsF);
    float* lf0 = A.f32(rowsF);
    float* fx = A.f32(rowsF * 512);
    Pl src = planes(A, rowsL, 24);
    // The oscillator's first two scan levels need f0 only: on short batches they run on the side branch from the start, beside
    // the frame-rate input conv and the SourceNet (joined in front of the third level, harmonic_source_cl).
    void* osc = A.bytes(osc_scratch_bytes(B, Lf));
    ARENA_OK();
    const bool scan_early = g_side_branch && !A.dry && side && rowsF <= 8192;
    if (scan_early) {
        TVC_CUDA(cudaEventRecord(fork, s));
        TVC_CUDA(cudaStreamWaitEvent(side, fork, 0));
        const int rc = osc_phase_scan(f0, osc, B, Lf, side);
        cudaEventRecord(scan_ev, side);                  // waited for in front of harmonic_source_cl (also after a failure:
        if (rc) { cudaStreamWaitEvent(s, scan_ev, 0); return rc; }          // a capture must end with every branch joined)
    }
    {
        const size_t m = A.mark();
        Pl cin = planes(A, rowsF, kFrameInCs);
        ARENA_OK();
        RUN(frame_prep(energy, f0, e_fr, lf0, B, Lf, s));
        RUN(cf_to_planes(content, cin.hi, cin.lo, B, kContent, Lf, kFrameInCs, TC_ACT_NONE, s, e_fr, lf0));
        CONV("tc_frame_in(", frame_in, ConvCall(cin, B, Lf).f32(fx, 512));
        A.release(m);
    }
    // ---- head start for the first Upsample block (decoder.py:174-176): its x2 resampler and c1 read the frame-rate product
    // only, not the down path.  On short batches they run on the side branch beside the SourceNet (whose kernels leave most SMs
    // idle); the block joins in front of c2, the first layer that needs the skip tensor.
    // (the buffers are reserved in sizing runs as well: they live from here to the block, on top of everything in between)
    const bool up0_bufs = g_side_branch && rowsF <= 8192 && kUpFac[0] * Lf < 384 && g_pad_max_t == 0;   // (< 384 rows: never a pruned level)
    const bool up0_early = up0_bufs && scan_early;
    float* xi_e = nullptr;
    Pl p0_e, p1_e;
    if (up0_bufs) {
        const int t0r = kUpFac[0] * Lf;
        xi_e = A.f32((long long)B * t0r * kUpCh[0]);
        p0_e = planes(A, (long long)B * t0r, kUpCh[0]);
        p1_e = planes(A, (long long)B * t0r, kUpCh[0]);
        ARENA_OK();
    }
    if (up0_early) {
        const int t0r = kUpFac[0] * Lf;
        TVC_CUDA(cudaEventRecord(fork, s));
        TVC_CUDA(cudaStreamWaitEvent(side, fork, 0));
        const bool pdl_was = t_pdl_suppress;
        t_pdl_suppress = true;
        int rc = interp_cl(fx + cm(0, 128, rowsF), B, Lf, t0r, (float)(1.0 / (double)kUpFac[0]), kUpCh[0], xi_e, nullptr, nullptr, p0_e.hi, p0_e.lo, side, 0);
        if (!rc) rc = tc_conv_k("tc_up0_c1(", up[0].c1, ConvCall(p0_e, B, t0r, 1).out(p1_e, TC_ACT_LRELU), side);
        t_pdl_suppress = pdl_was;
        cudaEventRecord(up0_ev, side);
        if (rc) { cudaStreamWaitEvent(s, up0_ev, 0); return rc; }
    }
    // ---- SourceNet (decoder.py:126-134) + dsp (decoder.py:259-266)
    {
        const size_t m = A.mark();
        Pl t1 = planes(A, rowsF, 128), xp = planes(A, rowsF, 128), t2p = planes(A, rowsF, 256);
        float* t2 = A.f32(rowsF * 256);
        float* hk = A.f32(rowsF * kHeadsCs);
        Pl yr = planes(A, rowsF, kBinsCs), yi = planes(A, rowsF, kBinsCs);
        float* cc = A.f32(rowsF * kBinsCs);
        float* ss = A.f32(rowsF * kBinsCs);
        float* noise = A.f32(rowsL);
        ARENA_OK();
        for (int i = 0; i < 3; ++i) {
            const Cnxt& c = mid[i];
            RUN(dwconv_ln_cl(fx, c.w7, c.wb, c.ln_g, c.ln_b, t1.hi, t1.lo, B, Lf, s));
            CONV("tc_cnxt_c2(", c.c2, ConvCall(t1, B, Lf).f32(t2, 256).epi(TC_ACT_GELU));
            RUN(grn_apply_cl(t2, c.grn_g, c.grn_b, t2p.hi, t2p.lo, B, 256, Lf, s));
            CONV("tc_cnxt_c3(", c.c3, ConvCall(t2p, B, Lf).res(fx, 512).f32(fx, 512).out(xp, TC_ACT_NONE));
        }
        CONV("tc_heads(", heads, ConvCall(xp, B, Lf).f32(hk, kHeadsCs).epi(TC_ACT_ELU1));
        RUN(noise_spectrum_cl(hk, rand01, rng_state, yr.hi, yr.lo, yi.hi, yi.lo, kBinsCs, B, Lf, s));
        // The cosine and sine products are independent.  On short batches each is one tile per CTA on fewer than half of the
        // SMs when the channel tiles are 128 wide, so they run side by side: the sine product on a forked branch, both
        // without programmatic dependent launch (a PDL successor would take the free SMs, tvc_common.cuh t_sm_cap).
        const long long idft_tiles_w = ((rowsF + 127) / 128) * dft_cos_w.n_tiles;
        if (g_side_branch && !A.dry && side && 2 * idft_tiles_w <= 148) {
            TVC_CUDA(cudaEventRecord(fork, s));
            TVC_CUDA(cudaStreamWaitEvent(side, fork, 0));
            const bool pdl_was = t_pdl_suppress;
            t_pdl_suppress = true;
            int rc = tc_conv_k("tc_idft(", dft_sin_w, ConvCall(yi, B, Lf).f32(ss, kBinsCs), side);
            if (!rc) rc = cudaEventRecord(join, side) != cudaSuccess;
            if (!rc) rc = tc_conv_k("tc_idft(", dft_cos_w, ConvCall(yr, B, Lf).f32(cc, kBinsCs), s);
            t_pdl_suppress = pdl_was;
            cudaStreamWaitEvent(s, join, 0);               // joined even after a failure: a capture must end on one stream
            TVC_TRY(rc);
        } else {
            CONV("tc_idft(", dft_cos, ConvCall(yr, B, Lf).f32(cc, kBinsCs));
            CONV("tc_idft(", dft_sin, ConvCall(yi, B, Lf).f32(ss, kBinsCs));
        }
        RUN(noise_ola_cl(cc, ss, noise, B, Lf, s));
        if (scan_early) TVC_CUDA(cudaStreamWaitEvent(s, scan_ev, 0));
        RUN(harmonic_source_cl(f0, hk + cm(0, kAmpsOff, rowsF), noise, energy, src.hi, src.lo, osc, B, Lf, s, scan_early));
        A.release(m);
    }
    // ---- FilterNet down path (decoder.py:206-213,225-229): skips at L, L/5, L/20, L/80, L/240
    // g_fuse_down: the 1/f resampler in front of every Downsample block (decoder.py:148) is evaluated by the epilogue of the conv
    // that produces its input (downs.0, then each block's c3), which writes the block's two operands directly: four launches
    // fewer, and the fp32 copy of the skip tensors (96 B per sample at the full rate) is never written.
    float* skip32[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    Pl skipP[5];
    int skipT[5];
    skipT[0] = L;
    const bool fuse_down = g_fuse_down;
    if (!fuse_down) skip32[0] = A.f32(rowsL * 24);
    skipP[0] = planes(A, rowsL, 24);
    // operands of the four Downsample blocks (raw planes for down_res, leaky-ReLU'd planes for c1, possibly padded)
    Pl xr[4], xa[4];
    int P1s[4];
    {
        int t = L;
        for (int i = 0; i < 4; ++i) {
            t /= kDownFac[i];
            skipT[i + 1] = t;
            P1s[i] = t <= g_pad_down_max_t ? 1 : 0;
            if (fuse_down) {
                xr[i] = planes(A, (long long)B * t, kDownIn[i]);
                xa[i] = planes(A, (long long)B * (t + 2 * P1s[i]), kDownIn[i]);
            }
        }
    }
    ARENA_OK();
    auto with_dec = [&](ConvCall c, int i) {      // attach Downsample block i's resampled operands to the producing conv
        const int fac = kDownFac[i];
        c.a.dec_f = fac; c.a.dec_pad = P1s[i]; c.a.dec_scale = (float)(1.0 / (1.0 / (double)fac));
        c.a.dec_r_hi = xr[i].hi; c.a.dec_r_lo = xr[i].lo; c.a.dec_a_hi = xa[i].hi; c.a.dec_a_lo = xa[i].lo;
        return c;
    };
    if (fuse_down) CONV("tc_down0(", down0, with_dec(ConvCall(src, B, L).out(skipP[0], TC_ACT_NONE), 0));
    else CONV("tc_down0(", down0, ConvCall(src, B, L).f32(skip32[0], 24).out(skipP[0], TC_ACT_NONE));
    for (int i = 0; i < 4; ++i) {
        const int cin = kDownIn[i], cout = kDownOut[i], fac = kDownFac[i];
        const int tin = skipT[i], tout = skipT[i + 1];    // exact: L = 480 * Lf
        const long long rows = (long long)B * tout;
        const bool last = i == 3;
        if (!fuse_down || last) skip32[i + 1] = A.f32(rows * cout);
        skipP[i + 1] = planes(A, rows, cout);
        const size_t m = A.mark();
        // stored replicate padding for short utterances: each tensor carries the dilation of the conv that reads it
        const bool pad = P1s[i] != 0;
        const int P1 = pad ? 1 : 0, P2 = pad ? 2 : 0, P4 = pad ? 4 : 0;
        if (!fuse_down) {
            xr[i] = planes(A, rows, cin);
            xa[i] = planes(A, (long long)B * (tout + 2 * P1), cin);
        }
        Pl a = planes(A, (long long)B * (tout + 2 * P2), cin), c = planes(A, (long long)B * (tout + 2 * P4), cin);
        ARENA_OK();
        if (!fuse_down) {
            const float scale = (float)(1.0 / (1.0 / (double)fac));   // F.interpolate(scale_factor=1/f)
            RUN(interp_cl(skip32[i], B, tin, tout, scale, cin, nullptr, xr[i].hi, xr[i].lo, xa[i].hi, xa[i].lo, s, P1));
        }
        const char* const n1[4] = {"tc_down1_c1(", "tc_down2_c1(", "tc_down3_c1(", "tc_down4_c1("};
        const char* const n2[4] = {"tc_down1_c2(", "tc_down2_c2(", "tc_down3_c2(", "tc_down4_c2("};
        const char* const n3[4] = {"tc_down1_c3(", "tc_down2_c3(", "tc_down3_c3(", "tc_down4_c3("};
        CONV(n1[i], down[i].c1, ConvCall(xa[i], B, tout, 1).pad(P1, P2).out(a, TC_ACT_LRELU));
        CONV(n2[i], down[i].c2, ConvCall(a, B, tout, 2).pad(P2, P4).out(c, TC_ACT_LRELU));
        if (fuse_down && !last)
            CONV(n3[i], down[i].c3, with_dec(ConvCall(c, B, tout, 4).pad(P4, 0).aux(xr[i]).out(skipP[i + 1], TC_ACT_NONE), i + 1));
        else
            CONV(n3[i], down[i].c3, ConvCall(c, B, tout, 4).pad(P4, 0).aux(xr[i]).f32(skip32[i + 1], cout).out(skipP[i + 1], TC_ACT_NONE));
        A.release(m);
    }
    // ---- FilterNet up path (decoder.py:214-219,230-233)
    // Output pruning (out_t0 / out_t1, fused block only): level i has to produce rows [wa[i], wb[i]) of its utterances only --
    // the rows the level above reads, widened by the 40 rows (1 + 3 + 9 + 27) over which a wrong edge value of the block's
    // dilated convs travels inward.  A pruned level works on compact tensors of wb - wa rows per utterance: the resampler
    // writes just those rows (coordinates on the full lengths), the FiLM condition is a row slice of the skip tensor, the
    // convs treat the window as an utterance (their replicate padding at a cut edge is wrong but stays inside the 40-row
    // margin), so every kept output sample is computed from exactly the values of the full run.  Levels whose window is
    // nearly everything, or too short to keep the conv kernels in the same (halo) tiling as the full run, are not pruned.
    int wa[5], wb[5];
    decoder_plan_windows(Lf, out_t0, out_t1, g_fused_up && g_prune_levels, wa, wb);
    const float* x = fx + cm(0, 128, rowsF);      // FilterNet x0 = channels [128, 512) of the frame-rate product
    int tin = Lf, tin_c = Lf, tin_off = 0;        // full length of the level input, rows it holds per utterance, first of them
    for (int i = 0; i < 5; ++i) {
        const Up& u = up[i];
        const int c = kUpCh[i], cn = kUpOut[i], fac = kUpFac[i];
        const int tfull = tin * fac;
        const int tout = wb[i] - wa[i];             // rows per utterance this level works on
        const long long rows = (long long)B * tout;
        const Pl& cond_full = skipP[4 - i];
        TVC_REQUIRE(skipT[4 - i] == tfull && cond_full.cs == c, "filter_net: skip %d shape mismatch", 4 - i);
        const Up& uc = up4_cat;
        // (workspace sizing runs this plan dry on a weightless model: the architecture's shapes are the supported ones)
        const bool fused = g_fused_up && i == 4 && fac == 5 && c == 24 &&
                           (A.dry || tc_up24_block_supported(uc.c1, uc.c2, uc.c3, uc.c4, uc.c5));
        if (fused) {
            // resampler, the five convs and the output layer in one kernel (tc_block.cu); same arithmetic as the launches below
            TcUpBlockArgs fa;
            fa.x4 = x; fa.c_hi = cond_full.hi; fa.c_lo = cond_full.lo; fa.out_w = out_w; fa.out_b = out_b; fa.out = out;
            fa.B = B; fa.T = tfull; fa.T4 = tin; fa.scale = (float)(1.0 / (double)fac);
            fa.t_lo = out_t0; fa.t_hi = out_t1;
            fa.x4_rows = tin_c; fa.x4_off = tin_off;
            if (!A.dry) {
                ProfScope ps("tc_up4_fused(", s);
                TVC_TRY(tc_up24_block_launch(uc.c1, uc.c2, uc.c3, uc.c4, uc.c5, fa, s));
            }
            A.release(m0);
            return 0;
        }
        TVC_REQUIRE(i < 4 || (tout == tfull && tin_c == tin), "filter_net: pruned levels need the fused block");
        const bool early = up0_early && i == 0, ebuf = up0_bufs && i == 0;
        float* xo = A.f32(rows * cn);
        const size_t m = A.mark();
        float* xi = ebuf ? xi_e : A.f32(rows * c);
        float* y = A.f32(rows * c);
        // stored replicate padding (see g_pad_max_t): p0 holds c1's input (1 row), later c3's (9); p1 c2's (3), later c4's (27)
        const bool windowed = tout != tfull || tin_c != tin;
        const bool pad = !fused && !windowed && tout <= g_pad_max_t;
        const int Q1 = pad ? 1 : 0, Q3 = pad ? 3 : 0, Q9 = pad ? 9 : 0, Q27 = pad ? 27 : 0;
        Pl p0 = ebuf ? p0_e : planes(A, (long long)B * (tout + 2 * Q9), c), p1 = ebuf ? p1_e : planes(A, (long long)B * (tout + 2 * Q27), c);
        Pl cond = cond_full;
        if (tout != tfull || A.dry) cond = planes(A, rows, c);      // (sizing runs reserve the slice: a pruned call then never needs more than the full plan)
        ARENA_OK();
        const float scale = (float)(1.0 / (double)fac);           // F.interpolate(scale_factor=f)
        if (early) {
            TVC_CUDA(cudaStreamWaitEvent(s, up0_ev, 0));      // resampler and c1 ran on the side branch
        } else if (windowed) {
            RUN(interp_cl(x, B, tin, tout, scale, c, xi, nullptr, nullptr, p0.hi, p0.lo, s, 0, tin_c, tin_off, wa[i]));
            if (tout != tfull) RUN(slice_planes_cl(cond_full.hi, cond_full.lo, cond.hi, cond.lo, B, tfull, c, wa[i], tout, s));
        } else {
            RUN(interp_cl(x, B, tin, tout, scale, c, xi, nullptr, nullptr, p0.hi, p0.lo, s, Q1));
        }
        const char* const un[5][5] = {{"tc_up0_c1(", "tc_up0_c2(", "tc_up0_c3(", "tc_up0_c4(", "tc_up0_c5("},
                                      {"tc_up1_c1(", "tc_up1_c2(", "tc_up1_c3(", "tc_up1_c4(", "tc_up1_c5("},
                                      {"tc_up2_c1(", "tc_up2_c2(", "tc_up2_c3(", "tc_up2_c4(", "tc_up2_c5("},
                                      {"tc_up3_c1(", "tc_up3_c2(", "tc_up3_c3(", "tc_up3_c4(", "tc_up3_c5("},
                                      {"tc_up4_c1(", "tc_up4_c2(", "tc_up4_c3(", "tc_up4_c4(", "tc_up4_c5("}};
        const bool has_w = i < 3 && !pad && !A.dry;
        const TcConvW& w1 = has_w ? pick_tiling(u.c1, up_w[i].c1, rows, 1) : u.c1;
        const TcConvW& w2 = has_w ? pick_tiling(u.c2, up_w[i].c2, rows, 3) : u.c2;
        const TcConvW& w3 = has_w ? pick_tiling(u.c3, up_w[i].c3, rows, 9) : u.c3;
        const TcConvW& w4 = has_w ? pick_tiling(u.c4, up_w[i].c4, rows, 27) : u.c4;
        if (!early) CONV(un[i][0], w1, ConvCall(p0, B, tout, 1).pad(Q1, Q3).out(p1, TC_ACT_LRELU));
        CONV(un[i][1], w2, ConvCall(p1, B, tout, 3).pad(Q3, Q9).aux(cond).res(xi, c).f32(y, c).out(p0, TC_ACT_LRELU));
        CONV(un[i][2], w3, ConvCall(p0, B, tout, 9).pad(Q9, Q27).out(p1, TC_ACT_LRELU));
        CONV(un[i][3], w4, ConvCall(p1, B, tout, 27).pad(Q27, 0).aux(cond).res(y, c).out(p0, TC_ACT_NONE));
        CONV(un[i][4], u.c5, ConvCall(p0, B, tout, 1).f32(xo, cn));
        A.release(m);
        x = xo; tin = tfull; tin_c = tout; tin_off = wa[i];
    }
    RUN(out_conv_k7_cl(x, out_w, out_b, out, B, L, s));
    A.release(m0);
    return 0;
}

// =============================================================================================
// Encoder on the tensor-core path
// =============================================================================================
EncoderTC::~EncoderTC() {
    in.free_all();
    for (Stack* st : {&ssl, &pitch}) {
        st->out.free_all(); st->out_w.free_all();
        for (auto& b : st->mid) { b.c2.free_all(); b.c2_w.free_all(); b.c3.free_all(); }
    }
    if (w7_buf) cudaFree(w7_buf);
}

int EncoderTC::init(const WeightStore& store) {
    HostW H;
    H.table = &store.table;
    H.flat.resize((size_t)store.table.total);
    TVC_CUDA(cudaMemcpy(H.flat.data(), store.flat, sizeof(float) * H.flat.size(), cudaMemcpyDeviceToHost));
    const int kSslDilTc[6] = {1, 3, 9, 1, 1, 1};               // encoder.py:80
    struct Def { Stack* st; const char* name; int C, layers, c0; };
    const Def defs[2] = {{&ssl, "ssl_feature_estimator", 384, 6, 0}, {&pitch, "pitch_estimator", 128, 4, 384}};
    // ---- merged input layers: rows [0, 384) ssl.input_layer, [384, 512) pitch.input_layer  (encoder.py:27,84)
    {
        std::vector<float> w((size_t)512 * kBins), b(512);
        for (const Def& d : defs) {
            const std::string p = d.name;
            const float *iw = H.get(p + ".input_layer.weight"), *ib = H.get(p + ".input_layer.bias");
            TVC_REQUIRE(iw && ib, "tc encoder weights: missing %s.input_layer", d.name);
            memcpy(&w[(size_t)d.c0 * kBins], iw, sizeof(float) * (size_t)d.C * kBins);
            memcpy(&b[d.c0], ib, sizeof(float) * d.C);
        }
        TVC_TRY(tc_pack_conv(w.data(), b.data(), 512, kBins, 1, nullptr, nullptr, 0, 0, 128, in));
    }
    // ---- depth-wise weights of all blocks, repacked [7][C]
    size_t w7_total = 0;
    for (const Def& d : defs) w7_total += (size_t)d.layers * 7 * d.C;
    std::vector<float> w7(w7_total);
    TVC_CUDA(cudaMalloc(&w7_buf, sizeof(float) * w7_total));
    size_t w7_off = 0;
    for (const Def& d : defs) {
        Stack& st = *d.st;
        const std::string p = d.name;
        st.C = d.C; st.c0 = d.c0;
        st.ln_g = store.raw(p + ".norm.gamma"); st.ln_b = store.raw(p + ".norm.beta");
        TVC_REQUIRE(st.ln_g && st.ln_b, "tc encoder weights: missing %s.norm", d.name);
        st.mid.resize(d.layers);
        for (int i = 0; i < d.layers; ++i) {
            const std::string q = p + ".mid_layers." + std::to_string(i);
            Blk& k = st.mid[i];
            k.dil = d.st == &ssl ? kSslDilTc[i] : 1;
            const float* dw = H.get(q + ".c1.weight");         // [C][1][7]
            TVC_REQUIRE(dw, "tc encoder weights: missing %s.c1", q.c_str());
            for (int c = 0; c < d.C; ++c)
                for (int j = 0; j < 7; ++j) w7[w7_off + (size_t)j * d.C + c] = dw[c * 7 + j];
            k.w7 = w7_buf + w7_off;
            w7_off += (size_t)7 * d.C;
            k.wb = store.raw(q + ".c1.bias");
            k.ln_g = store.raw(q + ".norm.gamma"); k.ln_b = store.raw(q + ".norm.beta");
            k.grn_g = store.raw(q + ".grn.gamma"); k.grn_b = store.raw(q + ".grn.beta");
            TVC_REQUIRE(k.wb && k.ln_g && k.ln_b && k.grn_g && k.grn_b, "tc encoder weights: incomplete %s", q.c_str());
            TVC_TRY(pack_named(H, q + ".c2", 128, k.c2));
            if (d.C == 384) TVC_TRY(pack_named(H, q + ".c2", 192, k.c2_w));
            TVC_TRY(pack_named(H, q + ".c3", d.C == 384 ? 96 : 64, k.c3));
        }
        TVC_TRY(pack_named(H, p + ".output_layer", 128, st.out));
        if (d.C == 384) TVC_TRY(pack_named(H, p + ".output_layer", 192, st.out_w));
    }
    TVC_CUDA(cudaMemcpy(w7_buf, w7.data(), sizeof(float) * w7_total, cudaMemcpyHostToDevice));
    ready = true;
    return 0;
}

int EncoderTC::forward(Arena& A, cudaStream_t s, const float* spec, float* z, float* logits, int B, int Lf) const {
    const long long rows = (long long)B * Lf;
    const size_t m0 = A.mark();
    constexpr int kSpecCs = 968;
    float* fx = A.f32(rows * 512);                 // [ssl x (384) | pitch x (128)], fp32 chunk-major
    {
        const size_t m = A.mark();
        Pl sp = planes(A, rows, kSpecCs);
        ARENA_OK();
        RUN(cf_to_planes(spec, sp.hi, sp.lo, B, kBins, Lf, kSpecCs, TC_ACT_NONE, s));
        CONV("tc_enc_in(", in, ConvCall(sp, B, Lf).f32(fx, 512));
        A.release(m);
    }
    const Stack* stacks[2] = {&ssl, &pitch};
    float* outs[2] = {z, logits};
    for (int k = 0; k < 2; ++k) {
        if (!outs[k]) continue;
        const Stack& st = *stacks[k];
        const int C = st.C, Cout = st.out.Cout;
        const size_t m = A.mark();
        float* x = fx + cm(0, st.c0, rows);        // this stack's channels of the merged product (a chunk-major view)
        Pl t1 = planes(A, rows, C), xp = planes(A, rows, C), t2p = planes(A, rows, 2 * C);
        float* t2 = A.f32(rows * 2 * C);
        float* y = A.f32(rows * Cout);
        ARENA_OK();
        RUN(cnxt_ln_cl(x, nullptr, nullptr, st.ln_g, st.ln_b, xp.hi, xp.lo, x, B, C, Lf, 1, s));       // encoder.py:29,86
        for (const Blk& b : st.mid) {              // convnext.py:49-58
            RUN(cnxt_ln_cl(x, b.w7, b.wb, b.ln_g, b.ln_b, t1.hi, t1.lo, nullptr, B, C, Lf, b.dil, s));
            CONV("tc_enc_c2(", A.dry ? b.c2 : pick_tiling(b.c2, b.c2_w, rows, 0), ConvCall(t1, B, Lf).f32(t2, 2 * C).epi(TC_ACT_GELU));
            RUN(grn_apply_cl(t2, b.grn_g, b.grn_b, t2p.hi, t2p.lo, B, 2 * C, Lf, s));
            CONV("tc_enc_c3(", b.c3, ConvCall(t2p, B, Lf).res(x, C).f32(x, C).out(xp, TC_ACT_NONE));
        }
        CONV("tc_enc_out(", A.dry ? st.out : pick_tiling(st.out, st.out_w, rows, 0), ConvCall(xp, B, Lf).f32(y, Cout));
        RUN(cl_to_cf(y, outs[k], B, Cout, Lf, Cout, s));
        A.release(m);
    }
    A.release(m0);
    return 0;
}

void set_fused_up(bool on) { g_fused_up = on; }
void set_fuse_down(bool on) { g_fuse_down = on; }
void set_prune_levels(bool on) { g_prune_levels = on; }
void set_side_branch(bool on) { g_side_branch = on; }
void set_wide_tiles(bool on) { g_wide_tiles = on; }
bool fused_up() { return g_fused_up; }
void set_pad_max_t(int up, int down) {
    if (up >= 0) g_pad_max_t = up > 2047 ? 2047 : up;
    if (down >= 0) g_pad_down_max_t = down > 2047 ? 2047 : down;
}
unsigned plan_options() {
    return (g_fused_up ? 1u : 0u) | ((unsigned)g_pad_max_t << 1) | ((unsigned)g_pad_down_max_t << 12) | (g_fuse_down ? 1u << 23 : 0u) | (g_prune_levels ? 1u << 24 : 0u) | (g_side_branch ? 1u << 25 : 0u) | (g_wide_tiles ? 1u << 26 : 0u);
}

}  // namespace tvc

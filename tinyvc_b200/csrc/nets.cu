// Parameter tables, weight packing and the launch sequences of the TinyVC networks.
#include "nets.cuh"
#include "nets_tc.cuh"

#include <cmath>
#include <cstring>

namespace tvc {

// =============================================================================================
// parameter tables (must reproduce torch's state_dict key order; checked from Python at load
// time against the module's own state_dict and in tests against the reference's key list)
// =============================================================================================
void ParamTable::add(const std::string& name, int d0, int d1, int d2) {
    ParamSpec s;
    s.name = name;
    s.d0 = d0; s.d1 = d1; s.d2 = d2;
    s.numel = (int64_t)d0 * (d1 ? d1 : 1) * (d2 ? d2 : 1);
    s.offset = total;
    total += s.numel;
    by_name[name] = (int)specs.size();
    specs.push_back(s);
}
void ParamTable::conv(const std::string& prefix, int cout, int cin, int k) {
    add(prefix + ".weight", cout, cin, k);
    add(prefix + ".bias", cout);
}
void ParamTable::convnext(const std::string& p, int c) {
    conv(p + ".c1", c, 1, 7);
    add(p + ".norm.gamma", c);
    add(p + ".norm.beta", c);
    conv(p + ".c2", 2 * c, c, 1);
    add(p + ".grn.beta", 1, 2 * c, 1);
    add(p + ".grn.gamma", 1, 2 * c, 1);
    conv(p + ".c3", c, 2 * c, 1);
}
const ParamSpec* ParamTable::find(const std::string& name) const {
    auto it = by_name.find(name);
    return it == by_name.end() ? nullptr : &specs[it->second];
}

static const int kUpCh[5] = {384, 192, 96, 48, 24};      // decoder.py:195
static const int kUpOut[5] = {192, 96, 48, 24, 24};
static const int kUpFac[5] = {2, 3, 4, 4, 5};            // decoder.py:196
static const int kDownIn[4] = {24, 48, 96, 192};
static const int kDownOut[4] = {48, 96, 192, 384};
static const int kDownFac[4] = {5, 4, 4, 3};
static const int kSslDil[6] = {1, 3, 9, 1, 1, 1};        // encoder.py:80

void build_decoder_table(ParamTable& t) {
    const std::string sn = "source_net", fn = "filter_net";
    t.conv(sn + ".content_in", 128, kContent, 1);
    t.conv(sn + ".energy_in", 128, 1, 1);
    t.conv(sn + ".f0_in", 128, 1, 1);
    for (int i = 0; i < 3; ++i) t.convnext(sn + ".mid_layers." + std::to_string(i), 128);
    t.conv(sn + ".to_amps", kOsc, 128, 1);
    t.conv(sn + ".to_kernel", kBins, 128, 1);
    t.conv(fn + ".content_in", 384, kContent, 1);
    t.conv(fn + ".f0_in", 384, 1, 1);
    t.conv(fn + ".downs.0", 24, kOsc + 2, 3);
    for (int i = 0; i < 4; ++i) {
        const std::string p = fn + ".downs." + std::to_string(i + 1);
        t.conv(p + ".down_res", kDownOut[i], kDownIn[i], 1);
        t.conv(p + ".c1", kDownIn[i], kDownIn[i], 3);
        t.conv(p + ".c2", kDownIn[i], kDownIn[i], 3);
        t.conv(p + ".c3", kDownOut[i], kDownIn[i], 3);
    }
    for (int i = 0; i < 5; ++i) {
        const std::string p = fn + ".ups." + std::to_string(i);
        const int c = kUpCh[i];
        t.conv(p + ".c1", c, c, 3);
        t.conv(p + ".c2", c, c, 3);
        t.conv(p + ".film1.to_shift", c, c, 1);
        t.conv(p + ".film1.to_scale", c, c, 1);
        t.conv(p + ".c3", c, c, 3);
        t.conv(p + ".c4", c, c, 3);
        t.conv(p + ".film2.to_shift", c, c, 1);
        t.conv(p + ".film2.to_scale", c, c, 1);
        t.conv(p + ".c5", kUpOut[i], c, 1);
    }
    t.conv(fn + ".output_layer", 1, 24, 7);
}

void build_encoder_table(ParamTable& t) {
    const std::string s = "ssl_feature_estimator", p = "pitch_estimator";
    t.conv(s + ".input_layer", 384, kBins, 1);
    t.add(s + ".norm.gamma", 384);
    t.add(s + ".norm.beta", 384);
    for (int i = 0; i < 6; ++i) t.convnext(s + ".mid_layers." + std::to_string(i), 384);
    t.conv(s + ".output_layer", kContent, 384, 1);
    t.conv(p + ".input_layer", 128, kBins, 1);
    t.add(p + ".norm.gamma", 128);
    t.add(p + ".norm.beta", 128);
    for (int i = 0; i < 4; ++i) t.convnext(p + ".mid_layers." + std::to_string(i), 128);
    t.conv(p + ".output_layer", 512, 128, 1);
}

// =============================================================================================
// weight store
// =============================================================================================
WeightStore::~WeightStore() {
    if (flat) cudaFree(flat);
    if (packed) cudaFree(packed);
}

int WeightStore::load(const float* params, int64_t numel) {
    TVC_REQUIRE(params != nullptr, "weights: null parameter pointer");
    TVC_REQUIRE(numel == table.total, "weights: got %lld parameters, the model has %lld", (long long)numel,
                (long long)table.total);
    TVC_CUDA(cudaMalloc(&flat, sizeof(float) * (size_t)table.total));
    TVC_CUDA(cudaMemcpy(flat, params, sizeof(float) * (size_t)table.total, cudaMemcpyDefault));
    packed_cap = table.total + table.total / 4 + (1 << 20);
    TVC_CUDA(cudaMalloc(&packed, sizeof(float) * (size_t)packed_cap));
    TVC_CUDA(cudaMemset(packed, 0, sizeof(float) * (size_t)packed_cap));
    packed_used = 0;
    return 0;
}

const float* WeightStore::raw(const std::string& name) const {
    const ParamSpec* s = table.find(name);
    return s ? flat + s->offset : nullptr;
}

float* WeightStore::take(int64_t n) {
    n = align_up(n, 16);
    if (packed_used + n > packed_cap) return nullptr;
    float* p = packed + packed_used;
    packed_used += n;
    return p;
}

int WeightStore::make_conv(const std::string& prefix, ConvW& out, cudaStream_t s) {
    const ParamSpec* w = table.find(prefix + ".weight");
    const ParamSpec* b = table.find(prefix + ".bias");
    TVC_REQUIRE(w && b, "weights: no conv named %s", prefix.c_str());
    out.Cout = w->d0; out.Cin = w->d1; out.K = w->d2;
    out.CoutP = (int)align_up(out.Cout, 4);
    float* dst = take((int64_t)out.K * out.Cin * out.CoutP);
    TVC_REQUIRE(dst, "weights: packed buffer exhausted at %s", prefix.c_str());
    TVC_TRY(repack_conv_weight(flat + w->offset, dst, out.Cout, out.Cin, out.K, out.CoutP, 0, s));
    out.w = dst;
    out.b = flat + b->offset;
    return 0;
}

// Two convs sharing an input, concatenated along Cout (FiLM: rows [0,C) to_scale, [C,2C) to_shift).
int WeightStore::make_conv_cat(const std::string& pa, const std::string& pb, ConvW& out, cudaStream_t s) {
    const ParamSpec *wa = table.find(pa + ".weight"), *ba = table.find(pa + ".bias");
    const ParamSpec *wb = table.find(pb + ".weight"), *bb = table.find(pb + ".bias");
    TVC_REQUIRE(wa && ba && wb && bb, "weights: no conv pair %s / %s", pa.c_str(), pb.c_str());
    TVC_REQUIRE(wa->d1 == wb->d1 && wa->d2 == wb->d2, "weights: %s / %s shapes differ", pa.c_str(), pb.c_str());
    out.Cout = wa->d0 + wb->d0; out.Cin = wa->d1; out.K = wa->d2;
    out.CoutP = (int)align_up(out.Cout, 4);
    float* dst = take((int64_t)out.K * out.Cin * out.CoutP);
    float* bias = take(out.Cout);
    TVC_REQUIRE(dst && bias, "weights: packed buffer exhausted at %s", pa.c_str());
    TVC_TRY(repack_conv_weight(flat + wa->offset, dst, wa->d0, out.Cin, out.K, out.CoutP, 0, s));
    TVC_TRY(repack_conv_weight(flat + wb->offset, dst, wb->d0, out.Cin, out.K, out.CoutP, wa->d0, s));
    TVC_CUDA(cudaMemcpyAsync(bias, flat + ba->offset, sizeof(float) * wa->d0, cudaMemcpyDeviceToDevice, s));
    TVC_CUDA(cudaMemcpyAsync(bias + wa->d0, flat + bb->offset, sizeof(float) * wb->d0, cudaMemcpyDeviceToDevice, s));
    out.w = dst;
    out.b = bias;
    return 0;
}

int WeightStore::make_cnxt(const std::string& p, int C, int dil, CnxtW& L, cudaStream_t s) {
    L.C = C; L.dil = dil;
    L.dw_w = raw(p + ".c1.weight"); L.dw_b = raw(p + ".c1.bias");
    L.ln_g = raw(p + ".norm.gamma"); L.ln_b = raw(p + ".norm.beta");
    L.grn_gamma = raw(p + ".grn.gamma"); L.grn_beta = raw(p + ".grn.beta");
    TVC_REQUIRE(L.dw_w && L.dw_b && L.ln_g && L.ln_b && L.grn_gamma && L.grn_beta, "weights: incomplete ConvNeXt %s", p.c_str());
    TVC_TRY(make_conv(p + ".c2", L.c2, s));
    TVC_TRY(make_conv(p + ".c3", L.c3, s));
    return 0;
}

// =============================================================================================
// arena
// =============================================================================================
void* Arena::bytes(size_t n) {
    n = (size_t)align_up((int64_t)n, 256);
    const size_t o = off;
    off += n;
    if (off > peak) peak = off;
    if (dry) return (void*)(uintptr_t)(4096 + o);   // never dereferenced
    if (off > cap) {
        overflow = true;
        return nullptr;
    }
    return base + o;
}

#define RUN(expr)                    \
    do {                             \
        if (!A.dry) {                \
            ProfScope ps__(#expr, s); \
            TVC_TRY(expr);           \
        }                            \
    } while (0)
#define ARENA_OK() TVC_REQUIRE(!A.overflow, "workspace too small: need at least %zu bytes, got %zu", A.peak, A.cap)

int conv_run(Arena& A, cudaStream_t s, const ConvW& W, const float* x, long long x_bs, float* y, long long y_bs, int B,
             int T, int dil, int pre, int epi, const float* res, long long res_bs, const float* film,
             long long film_bs, const float* pre_scale, const float* pre_shift) {
    if (A.dry) return 0;
    ConvParams p;
    p.x = x; p.x_bs = x_bs; p.w = W.w; p.bias = W.b; p.y = y; p.y_bs = y_bs;
    p.res = res; p.res_bs = res_bs; p.film = film; p.film_bs = film_bs;
    p.pre_scale = pre_scale; p.pre_shift = pre_shift;
    p.B = B; p.T = T; p.Cin = W.Cin; p.Cout = W.Cout; p.CoutP = W.CoutP; p.K = W.K; p.dil = dil;
    p.pre = pre; p.epi = epi;
    ProfScope ps(W.K == 1 ? "conv1d_k1(" : "conv1d_k3(", s);
    return conv1d_launch(p, s);
}

// ConvNeXt-v2 block (convnext.py:49-58), in place on x.  t1 [B,C,T], t2 [B,2C,T], sc [B,2C].
int convnext_forward(Arena& A, cudaStream_t s, const CnxtW& L, float* x, float* t1, float* t2, float* sc, int B, int T) {
    const int C = L.C;
    const long long bs1 = (long long)C * T, bs2 = 2LL * C * T;
    RUN(dwconv_ln(x, t1, L.dw_w, L.dw_b, L.ln_g, L.ln_b, B, C, T, L.dil, s));
    TVC_TRY(conv_run(A, s, L.c2, t1, bs1, t2, bs2, B, T, 1, PRE_NONE, EPI_GELU));
    RUN(grn_scale(t2, L.grn_gamma, sc, B, 2 * C, T, s));
    TVC_TRY(conv_run(A, s, L.c3, t2, bs2, x, bs1, B, T, 1, PRE_AFFINE, EPI_RES, x, bs1, nullptr, 0, sc, L.grn_beta));
    return 0;
}

// =============================================================================================
// decoder
// =============================================================================================
DecoderModel::~DecoderModel() {
    if (dft_buf) cudaFree(dft_buf);
    delete tc;
}

int DecoderModel::init(const float* params, int64_t numel) {
    build_decoder_table(store.table);
    TVC_TRY(store.load(params, numel));
    cudaStream_t s = 0;
    const std::string sn = "source_net", fn = "filter_net";
    TVC_TRY(store.make_conv(sn + ".content_in", sn_content_in, s));
    sn_energy_w = store.raw(sn + ".energy_in.weight"); sn_energy_b = store.raw(sn + ".energy_in.bias");
    sn_f0_w = store.raw(sn + ".f0_in.weight"); sn_f0_b = store.raw(sn + ".f0_in.bias");
    for (int i = 0; i < 3; ++i) TVC_TRY(store.make_cnxt(sn + ".mid_layers." + std::to_string(i), 128, 1, sn_mid[i], s));
    TVC_TRY(store.make_conv(sn + ".to_amps", sn_to_amps, s));
    TVC_TRY(store.make_conv(sn + ".to_kernel", sn_to_kernel, s));
    TVC_TRY(store.make_conv(fn + ".content_in", fn_content_in, s));
    fn_f0_w = store.raw(fn + ".f0_in.weight"); fn_f0_b = store.raw(fn + ".f0_in.bias");
    TVC_TRY(store.make_conv(fn + ".downs.0", fn_down0, s));
    for (int i = 0; i < 4; ++i) {
        const std::string p = fn + ".downs." + std::to_string(i + 1);
        fn_down[i].factor = kDownFac[i];
        TVC_TRY(store.make_conv(p + ".down_res", fn_down[i].res, s));
        TVC_TRY(store.make_conv(p + ".c1", fn_down[i].c1, s));
        TVC_TRY(store.make_conv(p + ".c2", fn_down[i].c2, s));
        TVC_TRY(store.make_conv(p + ".c3", fn_down[i].c3, s));
    }
    for (int i = 0; i < 5; ++i) {
        const std::string p = fn + ".ups." + std::to_string(i);
        Up& u = fn_up[i];
        u.factor = kUpFac[i];
        TVC_TRY(store.make_conv(p + ".c1", u.c1, s));
        TVC_TRY(store.make_conv(p + ".c2", u.c2, s));
        TVC_TRY(store.make_conv_cat(p + ".film1.to_scale", p + ".film1.to_shift", u.film1, s));
        TVC_TRY(store.make_conv(p + ".c3", u.c3, s));
        TVC_TRY(store.make_conv(p + ".c4", u.c4, s));
        TVC_TRY(store.make_conv_cat(p + ".film2.to_scale", p + ".film2.to_shift", u.film2, s));
        TVC_TRY(store.make_conv(p + ".c5", u.c5, s));
    }
    fn_out_w = store.raw(fn + ".output_layer.weight"); fn_out_b = store.raw(fn + ".output_layer.bias");
    TVC_REQUIRE(sn_energy_w && sn_f0_w && fn_f0_w && fn_out_w, "decoder: missing parameters");

    // inverse real-DFT basis, fp64 on the host then rounded once:
    //   cosB[f][p] = w_f cos(2 pi f p / N) / N,  sinB[f][p] = w_f sin(2 pi f p / N) / N,  w_0 = w_{N/2} = 1, else 2
    // stored as packed 1x1-conv weights [Cin = f][CoutP = 964 >= p].
    const int P = (int)align_up(kBins, 4);
    std::vector<float> host((size_t)2 * kBins * P, 0.f);
    for (int f = 0; f < kBins; ++f) {
        const double wf = (f == 0 || f == kNfft / 2) ? 1.0 : 2.0;
        for (int p = 0; p < kBins; ++p) {
            const int r = (int)(((long long)f * p) % kNfft);     // exact argument reduction
            const double ang = 2.0 * M_PI * (double)r / (double)kNfft;
            host[(size_t)f * P + p] = (float)(wf * std::cos(ang) / (double)kNfft);
            host[(size_t)kBins * P + (size_t)f * P + p] = (float)(wf * std::sin(ang) / (double)kNfft);
        }
    }
    TVC_CUDA(cudaMalloc(&dft_buf, sizeof(float) * host.size()));
    TVC_CUDA(cudaMemcpy(dft_buf, host.data(), sizeof(float) * host.size(), cudaMemcpyHostToDevice));
    dft_cos.w = dft_buf; dft_cos.b = nullptr; dft_cos.Cin = kBins; dft_cos.Cout = kBins; dft_cos.CoutP = P; dft_cos.K = 1;
    dft_sin = dft_cos;
    dft_sin.w = dft_buf + (size_t)kBins * P;
    TVC_CUDA(cudaStreamSynchronize(s));
    tc = new DecoderTC();
    TVC_TRY(tc->init(store));
    return 0;
}

// SourceNet.forward (decoder.py:126-134).  e_fr/lf0 come from frame_prep.
int DecoderModel::source_net(Arena& A, cudaStream_t s, const float* content, const float* e_fr, const float* lf0,
                             float* amps, float* kern, int B, int Lf) {
    const int C = 128;
    const size_t m = A.mark();
    float* x = A.f32((int64_t)B * C * Lf);
    float* t1 = A.f32((int64_t)B * C * Lf);
    float* t2 = A.f32((int64_t)B * 2 * C * Lf);
    float* sc = A.f32((int64_t)B * 2 * C);
    ARENA_OK();
    const long long bs = (long long)C * Lf;
    TVC_TRY(conv_run(A, s, sn_content_in, content, (long long)kContent * Lf, x, bs, B, Lf, 1, PRE_NONE, EPI_NONE));
    RUN(rank1_add(x, e_fr, sn_energy_w, sn_energy_b, lf0, sn_f0_w, sn_f0_b, B, C, Lf, s));
    for (int i = 0; i < 3; ++i) TVC_TRY(convnext_forward(A, s, sn_mid[i], x, t1, t2, sc, B, Lf));
    TVC_TRY(conv_run(A, s, sn_to_amps, x, bs, amps, (long long)kOsc * Lf, B, Lf, 1, PRE_NONE, EPI_ELU1));
    TVC_TRY(conv_run(A, s, sn_to_kernel, x, bs, kern, (long long)kBins * Lf, B, Lf, 1, PRE_NONE, EPI_ELU1));
    A.release(m);
    return 0;
}

// Decoder.dsp (decoder.py:259-266): channels [0,15) of src = harmonics * interp(amps), channel 15 = noise.
int DecoderModel::dsp(Arena& A, cudaStream_t s, const float* f0, const float* amps, const float* kern,
                      const float* rand01, float* src, long long src_bs, int B, int Lf) {
    const size_t m = A.mark();
    float* yri = A.f32((int64_t)B * 2 * kBins * Lf);
    float* cs = A.f32((int64_t)B * 2 * kBins * Lf);
    ARENA_OK();
    const long long bs = 2LL * kBins * Lf;
    RUN(harmonic_osc(f0, amps, src, src_bs, B, Lf, s));
    RUN(noise_spectrum(kern, rand01, yri, B, Lf, s));
    TVC_TRY(conv_run(A, s, dft_cos, yri, bs, cs, bs, B, Lf, 1, PRE_NONE, EPI_NONE));
    TVC_TRY(conv_run(A, s, dft_sin, yri + (long long)kBins * Lf, bs, cs + (long long)kBins * Lf, bs, B, Lf, 1, PRE_NONE, EPI_NONE));
    RUN(noise_ola(cs, src, src_bs, kOsc, B, Lf, s));
    A.release(m);
    return 0;
}

// FilterNet.forward (decoder.py:222-233).  src17 = cat(source16, energy) [B,17,L].
int DecoderModel::filter_net(Arena& A, cudaStream_t s, const float* content, const float* lf0, const float* src17,
                             float* out, int B, int Lf) {
    const int L = Lf * kFrame;
    const size_t m0 = A.mark();
    // ---- down path: skips at L, L/5, L/20, L/80, L/240 ----
    float* skip[5];
    int skipT[5], skipC[5];
    skipT[0] = L; skipC[0] = 24;
    skip[0] = A.f32((int64_t)B * 24 * L);
    ARENA_OK();
    TVC_TRY(conv_run(A, s, fn_down0, src17, 17LL * L, skip[0], 24LL * L, B, L, 1, PRE_NONE, EPI_NONE));
    for (int i = 0; i < 4; ++i) {
        const Down& d = fn_down[i];
        const int cin = kDownIn[i], cout = kDownOut[i];
        const int fac = kDownFac[i];
        const int tin = skipT[i], tout = tin / fac;     // exact: L = 480*Lf
        skipT[i + 1] = tout; skipC[i + 1] = cout;
        skip[i + 1] = A.f32((int64_t)B * cout * tout);
        const size_t m = A.mark();
        float* xi = A.f32((int64_t)B * cin * tout);
        float* r = A.f32((int64_t)B * cout * tout);
        float* a = A.f32((int64_t)B * cin * tout);
        float* c = A.f32((int64_t)B * cin * tout);
        ARENA_OK();
        const long long bi = (long long)cin * tout, bo = (long long)cout * tout;
        // F.interpolate(scale_factor=1/f): coordinate scale = float(1/(1/f))
        const float scale = (float)(1.0 / (1.0 / (double)fac));
        RUN(interp_linear(skip[i], xi, (long long)B * cin, tin, tout, scale, s));
        TVC_TRY(conv_run(A, s, d.res, xi, bi, r, bo, B, tout, 1, PRE_NONE, EPI_NONE));
        TVC_TRY(conv_run(A, s, d.c1, xi, bi, a, bi, B, tout, 1, PRE_LRELU, EPI_NONE));
        TVC_TRY(conv_run(A, s, d.c2, a, bi, c, bi, B, tout, 2, PRE_LRELU, EPI_NONE));
        TVC_TRY(conv_run(A, s, d.c3, c, bi, skip[i + 1], bo, B, tout, 4, PRE_LRELU, EPI_RES, r, bo));
        A.release(m);
    }
    // ---- frame-rate conditioning: content_in(content) + f0_in(log f0) ----
    float* x = A.f32((int64_t)B * 384 * Lf);
    ARENA_OK();
    TVC_TRY(conv_run(A, s, fn_content_in, content, (long long)kContent * Lf, x, 384LL * Lf, B, Lf, 1, PRE_NONE, EPI_NONE));
    RUN(rank1_add(x, nullptr, nullptr, nullptr, lf0, fn_f0_w, fn_f0_b, B, 384, Lf, s));
    // ---- up path ----
    int tin = Lf;
    for (int i = 0; i < 5; ++i) {
        const Up& u = fn_up[i];
        const int c = kUpCh[i], cn = kUpOut[i];
        const int fac = kUpFac[i];
        const int tout = tin * fac;
        const float* cond = skip[4 - i];
        TVC_REQUIRE(skipT[4 - i] == tout && skipC[4 - i] == c, "filter_net: skip %d shape mismatch", 4 - i);
        float* xo = A.f32((int64_t)B * cn * tout);
        const size_t m = A.mark();
        float* xi = A.f32((int64_t)B * c * tout);
        float* a = A.f32((int64_t)B * c * tout);
        float* y = A.f32((int64_t)B * c * tout);
        float* ss = A.f32((int64_t)B * 2 * c * tout);
        ARENA_OK();
        const long long bc = (long long)c * tout, b2 = 2LL * c * tout;
        const float scale = (float)(1.0 / (double)fac);    // F.interpolate(scale_factor=f)
        RUN(interp_linear(x, xi, (long long)B * c, tin, tout, scale, s));
        TVC_TRY(conv_run(A, s, u.c1, xi, bc, a, bc, B, tout, 1, PRE_LRELU, EPI_NONE));
        TVC_TRY(conv_run(A, s, u.film1, cond, bc, ss, b2, B, tout, 1, PRE_NONE, EPI_NONE));
        TVC_TRY(conv_run(A, s, u.c2, a, bc, y, bc, B, tout, 3, PRE_LRELU, EPI_FILM_RES, xi, bc, ss, b2));
        TVC_TRY(conv_run(A, s, u.c3, y, bc, a, bc, B, tout, 9, PRE_LRELU, EPI_NONE));
        TVC_TRY(conv_run(A, s, u.film2, cond, bc, ss, b2, B, tout, 1, PRE_NONE, EPI_NONE));
        TVC_TRY(conv_run(A, s, u.c4, a, bc, xi, bc, B, tout, 27, PRE_LRELU, EPI_FILM_RES, y, bc, ss, b2));
        TVC_TRY(conv_run(A, s, u.c5, xi, bc, xo, (long long)cn * tout, B, tout, 1, PRE_NONE, EPI_NONE));
        A.release(m);
        x = xo;
        tin = tout;
    }
    RUN(out_conv_k7(x, fn_out_w, fn_out_b, out, B, 24, L, s));
    A.release(m0);
    return 0;
}

// Decoder.infer (decoder.py:253-257).
int DecoderModel::infer(Arena& A, cudaStream_t s, const float* content, const float* f0, const float* energy,
                        const float* rand01, float* out, int B, int Lf, int impl, int out_t0, int out_t1) {
    if (impl == CONV_IMPL_TC) {
        static const DecoderTC shape_only;   // dry runs (workspace sizing) never touch weights
        TVC_REQUIRE(A.dry || (tc && tc->ready), "decoder: tensor-core plan not initialised");
        return (A.dry && !tc ? shape_only : *tc).infer(A, s, content, f0, energy, rand01, out, B, Lf, out_t0, out_t1);
    }
    const int L = Lf * kFrame;
    const size_t m = A.mark();
    float* e_fr = A.f32((int64_t)B * Lf);
    float* lf0 = A.f32((int64_t)B * Lf);
    float* amps = A.f32((int64_t)B * kOsc * Lf);
    float* kern = A.f32((int64_t)B * kBins * Lf);
    float* src17 = A.f32((int64_t)B * 17 * L);
    ARENA_OK();
    RUN(frame_prep(energy, f0, e_fr, lf0, B, Lf, s));
    TVC_TRY(source_net(A, s, content, e_fr, lf0, amps, kern, B, Lf));
    TVC_TRY(dsp(A, s, f0, amps, kern, rand01, src17, 17LL * L, B, Lf));
    if (!A.dry)   // torch.cat([source, energy], dim=1)  (decoder.py:224): energy -> channel 16
        TVC_CUDA(cudaMemcpy2DAsync(src17 + 16LL * L, sizeof(float) * 17 * L, energy, sizeof(float) * L,
                                   sizeof(float) * L, B, cudaMemcpyDeviceToDevice, s));
    TVC_TRY(filter_net(A, s, content, lf0, src17, out, B, Lf));
    A.release(m);
    return 0;
}

// =============================================================================================
// encoder
// =============================================================================================
EncoderModel::~EncoderModel() { delete tc; }

int EncoderModel::init(const float* params, int64_t numel) {
    build_encoder_table(store.table);
    TVC_TRY(store.load(params, numel));
    cudaStream_t s = 0;
    struct Def { Stack* st; const char* name; int C; int layers; };
    const Def defs[2] = {{&ssl, "ssl_feature_estimator", 384, 6}, {&pitch, "pitch_estimator", 128, 4}};
    for (const Def& d : defs) {
        const std::string p = d.name;
        d.st->C = d.C;
        TVC_TRY(store.make_conv(p + ".input_layer", d.st->in, s));
        d.st->ln_g = store.raw(p + ".norm.gamma");
        d.st->ln_b = store.raw(p + ".norm.beta");
        TVC_REQUIRE(d.st->ln_g && d.st->ln_b, "encoder: missing %s.norm", d.name);
        d.st->mid.resize(d.layers);
        for (int i = 0; i < d.layers; ++i) {
            const int dil = (d.st == &ssl) ? kSslDil[i] : 1;
            TVC_TRY(store.make_cnxt(p + ".mid_layers." + std::to_string(i), d.C, dil, d.st->mid[i], s));
        }
        TVC_TRY(store.make_conv(p + ".output_layer", d.st->out, s));
    }
    TVC_CUDA(cudaStreamSynchronize(s));
    tc = new EncoderTC();
    TVC_TRY(tc->init(store));
    return 0;
}

// SSLFeatureEstimator.forward / PitchEstimator.forward (encoder.py:33-38, 89-94): out [B, Cout, Lf].
int EncoderModel::run_stack(Arena& A, cudaStream_t s, const Stack& st, const float* spec, float* out, int B, int Lf) {
    const int C = st.C;
    const size_t m = A.mark();
    float* x = A.f32((int64_t)B * C * Lf);
    float* t1 = A.f32((int64_t)B * C * Lf);
    float* t2 = A.f32((int64_t)B * 2 * C * Lf);
    float* sc = A.f32((int64_t)B * 2 * C);
    ARENA_OK();
    const long long bs = (long long)C * Lf;
    TVC_TRY(conv_run(A, s, st.in, spec, (long long)kBins * Lf, x, bs, B, Lf, 1, PRE_NONE, EPI_NONE));
    RUN(dwconv_ln(x, x, nullptr, nullptr, st.ln_g, st.ln_b, B, C, Lf, 1, s));
    for (const CnxtW& L : st.mid) TVC_TRY(convnext_forward(A, s, L, x, t1, t2, sc, B, Lf));
    TVC_TRY(conv_run(A, s, st.out, x, bs, out, (long long)st.out.Cout * Lf, B, Lf, 1, PRE_NONE, EPI_NONE));
    A.release(m);
    return 0;
}

}  // namespace tvc

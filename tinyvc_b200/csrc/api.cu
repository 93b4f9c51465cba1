// extern "C" entry points declared in include/tinyvc_b200.h.
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <atomic>
#include <map>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only: ranges cost nothing unless a profiler is attached

#include "../../include/tinyvc_b200.h"
#include "nets.cuh"
#include "nets_tc.cuh"
#include "tc_kernels.cuh"
#include "tc_conv.cuh"

namespace tvc {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static bool g_use_graphs = true;
static bool g_encoder_tc = true;     // tvc_set_option("encoder_impl", "tc"|"fp32"): Encoder on tcgen05 (default) or exact-fp32 CUDA cores
// tvc_set_option("pitch_impl", "fp32"|"tc"): the PitchEstimator stack (a ninth of the Encoder's MACs) stays on exact fp32 by
// default even when the content stack runs on tensor cores: the decoder integrates f0 into the oscillator phase over the whole
// utterance, which amplifies an f0 error ~5 000 x into the waveform (DESIGN.md "conditioning of the reference"), while the
// content vector only has to keep the kNN ranking.
static bool g_pitch_tc = false;
// tvc_set_option("encoder_share", "0"|"1"): on short batches cap the content stack's persistent kernels at kEncoderShareSms SMs
// (and launch them without PDL) so that the fp32 pitch stack on the side stream gets SMs of its own.
static bool g_encoder_share = true;
// Measured (profiles/r02ah_encoder_share_ab.log): 128 streams x 28 frames 0.913 -> 0.687 ms, config 3 (51 200 frames) 7.74 -> 6.10 ms;
// a single stream (28 frames) is 3 % slower with it, so batches under kEncoderShareMinFrames keep PDL and every SM.
static bool g_encoder_share_force = false;
static int g_encoder_share_sms = 0;                         // ("encoder_share", "<n>") forces this cap at every size (A/B runs)
constexpr long long kEncoderShareMinFrames = 64, kEncoderShareFrames = 8192;
constexpr int kEncoderShareSmsShort = 112, kEncoderShareSmsLong = 136;  // short batches are latency-bound in both stacks; long ones need the SMs on the tensor-core stack
static int g_probe_pad_in = 0, g_probe_pad_out = 0;     // tvc_set_option("probe_pad", ...): tests only
thread_local int t_sm_cap = 0;
thread_local bool t_pdl_suppress = false;
bool g_pdl = true;       // programmatic dependent launch between the decoder plan's kernels: the next kernel's CTAs start their
                         // set-up (mbarriers, TMEM, weight prefetch) on SMs the current one leaves idle (most layers of the
                         // low rates have fewer tiles than SMs); same-box A/B 0.986 -> 0.955 ms (profiles/r02h_pdl_ab.log)

// ---- event profiler ----------------------------------------------------------------------------
struct ProfEntry { std::string name; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfEntry> g_prof;

static bool g_nvtx = false;          // tvc_set_option("nvtx", "1"): an NVTX range around every launcher of the plans (nsys / ncu --nvtx)
ProfScope::ProfScope(const char* name, cudaStream_t s) : stream(s) {
    if (g_nvtx) {
        nvtxRangePushA(name);
        nvtx = true;
    }
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    ProfEntry e;
    const char* paren = strchr(name, '(');
    e.name = paren ? std::string(name, paren - name) : std::string(name);
    if (cudaEventCreate(&e.a) != cudaSuccess || cudaEventCreate(&e.b) != cudaSuccess) return;
    cudaEventRecord(e.a, s);
    slot = (int)g_prof.size();
    g_prof.push_back(e);
}
ProfScope::~ProfScope() {
    if (nvtx) nvtxRangePop();
    if (slot < 0) return;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    cudaEventRecord(g_prof[slot].b, stream);
}

// Kernel attributes are per device: the set-up runs the first time each device is seen (every entry point that launches
// the conv kernels goes through here, so `Decoder().to("cuda:1")` after cuda:0 works in one process).
static int global_init() {
    static PerDeviceOnce once;
    return once.run([] {
        TVC_TRY(conv1d_init());
        return tc_conv_init();
    });
}

static ParamTable& table_of(int kind) {
    static ParamTable dec, enc;
    static std::once_flag once;
    std::call_once(once, [] {
        build_decoder_table(dec);
        build_encoder_table(enc);
    });
    return kind == 0 ? dec : enc;
}

struct IndexModel {
    float* index_w = nullptr;    // [C][NP] (normalised for cos)
    float* index_nc = nullptr;   // [N][C] raw
    float* bias = nullptr;       // [N]
    int N = 0, NP = 0, metric = 0;
    ConvW as_conv;
    // tensor-core screening (cos metric): the normalised index as a packed 1x1 conv with Cout = N, and a row-major
    // copy of the same normalised values for the exact re-scoring of the candidates (knn.cu)
    TcConvW tc;
    float* index_wn = nullptr;   // [N][C]
    bool screened = false;
    ~IndexModel() {
        tc.free_all();
        if (index_wn) cudaFree(index_wn);
        if (index_w) cudaFree(index_w);
        if (index_nc) cudaFree(index_nc);
        if (bias) cudaFree(bias);
    }
};

}  // namespace tvc

using namespace tvc;

// A captured CUDA graph of one Decoder.infer launch sequence, valid for one exact set of buffers.
struct DecoderGraphKey {
    const void *content, *f0, *energy, *rand01, *out, *ws;
    int B, Lf, impl;
    unsigned opts;                       // plan options (nets_tc.cuh plan_options): they change the launch sequence
    int t0, t1;                          // output range (tvc_decoder_infer_range)
    bool operator==(const DecoderGraphKey& o) const {
        return content == o.content && f0 == o.f0 && energy == o.energy && rand01 == o.rand01 && out == o.out &&
               ws == o.ws && B == o.B && Lf == o.Lf && impl == o.impl && opts == o.opts && t0 == o.t0 && t1 == o.t1;
    }
};
struct DecoderGraph {
    DecoderGraphKey key;
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0, last_use = 0;
};
struct tvc_decoder {
    DecoderModel m;
    std::mutex mu;
    std::vector<DecoderGraph> graphs;          // small LRU cache
    std::vector<DecoderGraphKey> seen;         // buffer sets seen once (a second sighting triggers capture)
    cudaStream_t cap_stream = nullptr;
    unsigned long long tick = 0;
    ~tvc_decoder() {
        for (DecoderGraph& g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        if (cap_stream) cudaStreamDestroy(cap_stream);
    }
};
struct tvc_encoder {
    EncoderModel m;
    // the fp32 pitch stack runs beside the tensor-core content stack on a side stream (fork / join by events)
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    ~tvc_encoder() {
        if (fork) cudaEventDestroy(fork);
        if (join) cudaEventDestroy(join);
        if (side) cudaStreamDestroy(side);
    }
};
struct tvc_index { IndexModel m; };

#define API_BEGIN try {
#define API_END                                              \
    }                                                        \
    catch (const std::exception& e) {                        \
        tvc::set_error("exception: %s", e.what());           \
        return 3;                                            \
    }                                                        \
    catch (...) {                                            \
        tvc::set_error("unknown exception");                 \
        return 3;                                            \
    }

#pragma GCC visibility push(default)
extern "C" {

const char* tvc_last_error(void) { return g_err; }
const char* tvc_version(void) { return "tinyvc_b200 0.1 (sm_100a)"; }

int tvc_set_option(const char* key, const char* value) {
    if (!key || !value) return 2;
    if (!strcmp(key, "conv_impl")) {
        if (!strcmp(value, "fp32")) { g_conv_impl = CONV_IMPL_FP32; return 0; }
        if (!strcmp(value, "tc")) { g_conv_impl = CONV_IMPL_TC; return 0; }
        set_error("conv_impl: unknown value '%s'", value);
        return 2;
    }
    if (!strcmp(key, "encoder_impl")) {
        if (!strcmp(value, "fp32")) { g_encoder_tc = false; return 0; }
        if (!strcmp(value, "tc")) { g_encoder_tc = true; return 0; }
        set_error("encoder_impl: unknown value '%s'", value);
        return 2;
    }
    if (!strcmp(key, "nvtx")) { g_nvtx = !strcmp(value, "1"); return 0; }
    if (!strcmp(key, "encoder_share")) {
        const int n = atoi(value);
        g_encoder_share = n != 0;
        g_encoder_share_force = n > 1;                      // an explicit cap shares at every size (A/B runs)
        if (n > 1) g_encoder_share_sms = n;
        return 0;
    }
    if (!strcmp(key, "pitch_impl")) {
        if (!strcmp(value, "fp32")) { g_pitch_tc = false; return 0; }
        if (!strcmp(value, "tc")) { g_pitch_tc = true; return 0; }
        set_error("pitch_impl: unknown value '%s'", value);
        return 2;
    }
    if (!strcmp(key, "tc_trace")) return tc_trace_arm(value);          // developer: "k0,k1,..." launch ordinals
    if (!strcmp(key, "tc_trace_dump")) return tc_trace_dump(value);    // developer: write the timeline to a file
    if (!strcmp(key, "fused_up")) { set_fused_up(!strcmp(value, "1")); return 0; }
    if (!strcmp(key, "fuse_down")) { set_fuse_down(!strcmp(value, "1")); return 0; }
    if (!strcmp(key, "prune_levels")) { set_prune_levels(!strcmp(value, "1")); return 0; }
    if (!strcmp(key, "side_branch") || !strcmp(key, "idft_pair")) { set_side_branch(!strcmp(value, "1")); return 0; }   // (idft_pair: the name in the A/B logs)
    if (!strcmp(key, "wide_tiles")) { set_wide_tiles(!strcmp(value, "1")); return 0; }
    if (!strcmp(key, "pad_up_max_t")) { set_pad_max_t(atoi(value), -1); return 0; }
    if (!strcmp(key, "pad_down_max_t")) { set_pad_max_t(-1, atoi(value)); return 0; }
    if (!strcmp(key, "probe_pad")) {                                   // tests: tvc_tc_conv_probe in padded mode
        int a = 0, b = 0;
        if (sscanf(value, "%d,%d", &a, &b) != 2 || a < 0 || b < 0) return 1;
        g_probe_pad_in = a; g_probe_pad_out = b;
        return 0;
    }
    if (!strcmp(key, "pdl")) {
        g_pdl = !strcmp(value, "1");
        return 0;
    }
    if (!strcmp(key, "graphs")) {
        g_use_graphs = !strcmp(value, "1");
        return 0;
    }
    if (!strcmp(key, "profile")) {
        std::lock_guard<std::mutex> lock(g_prof_mu);
        g_prof_on = !strcmp(value, "1");
        return 0;
    }
    set_error("unknown option '%s'", key);
    return 2;
}

unsigned long long tvc_launch_count(void) { return g_launches.load(); }

int tvc_profile_report(char* buf, size_t n) {
    API_BEGIN
    TVC_REQUIRE(buf && n > 2, "tvc_profile_report: need a buffer");
    TVC_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> acc;
    for (ProfEntry& e : g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            acc[e.name].first += 1;
            acc[e.name].second += ms;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    g_prof.clear();
    std::string js = "{";
    for (auto& kv : acc) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", js.size() > 1 ? ", " : "",
                 kv.first.c_str(), kv.second.first, kv.second.second);
        js += tmp;
    }
    js += "}";
    TVC_REQUIRE(js.size() + 1 <= n, "tvc_profile_report: buffer of %zu bytes too small (%zu needed)", n, js.size() + 1);
    memcpy(buf, js.c_str(), js.size() + 1);
    return 0;
    API_END
}

int tvc_measure_fp32_peak(double* tflops, void* stream) {
    API_BEGIN
    TVC_REQUIRE(tflops, "tvc_measure_fp32_peak: null argument");
    return measure_fp32_peak(tflops, (cudaStream_t)stream);
    API_END
}

int tvc_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]) {
    API_BEGIN
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    TVC_REQUIRE(ptr && handle && bytes > 0, "tvc_peer_alloc: bad arguments");
    void* p = nullptr;
    TVC_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        TVC_CUDA(e);
    }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return 0;
    API_END
}
int tvc_peer_open(const unsigned char handle[64], void** ptr) {
    API_BEGIN
    TVC_REQUIRE(ptr && handle, "tvc_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    TVC_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = p;
    return 0;
    API_END
}
int tvc_peer_close(void* ptr) {
    API_BEGIN
    if (ptr) TVC_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
    API_END
}
int tvc_peer_free(void* ptr) {
    API_BEGIN
    if (ptr) TVC_CUDA(cudaFree(ptr));
    return 0;
    API_END
}

int tvc_param_count(int kind) { return (kind == 0 || kind == 1) ? (int)table_of(kind).specs.size() : -1; }
const char* tvc_param_name(int kind, int i) {
    if (kind != 0 && kind != 1) return nullptr;
    ParamTable& t = table_of(kind);
    return (i >= 0 && i < (int)t.specs.size()) ? t.specs[i].name.c_str() : nullptr;
}
int64_t tvc_param_numel(int kind, int i) {
    if (kind != 0 && kind != 1) return -1;
    ParamTable& t = table_of(kind);
    return (i >= 0 && i < (int)t.specs.size()) ? t.specs[i].numel : -1;
}
int64_t tvc_param_total(int kind) { return (kind == 0 || kind == 1) ? table_of(kind).total : -1; }

// ---------------------------------------------------------------------------------------------- decoder
int tvc_decoder_create(const float* params, int64_t numel, tvc_decoder_t* out) {
    API_BEGIN
    TVC_REQUIRE(out, "tvc_decoder_create: null out pointer");
    TVC_TRY(global_init());
    tvc_decoder* h = new (std::nothrow) tvc_decoder();
    TVC_REQUIRE(h, "tvc_decoder_create: out of host memory");
    const int r = h->m.init(params, numel);
    if (r) { delete h; return r; }
    *out = h;
    return 0;
    API_END
}
int tvc_decoder_destroy(tvc_decoder_t h) { delete h; return 0; }

static void shapes_ok_msg(int B, int Lf) { set_error("invalid shape B=%d Lf=%d", B, Lf); }
#define CHECK_SHAPES() do { if (B <= 0 || Lf <= 0 || (long long)B * Lf * 480 > (1LL << 31) - 1) { shapes_ok_msg(B, Lf); return 2; } } while (0)

size_t tvc_decoder_workspace_bytes(int B, int Lf) {
    if (B <= 0 || Lf <= 0) return 0;
    static DecoderModel shape_only;   // dry runs never touch weights (nor any mutable state of the model)
    // the larger of the two execution plans, so the caller's buffer fits every entry point (tvc_source_net / tvc_dsp /
    // tvc_filter_net always run the exact-fp32 plan) whichever options are active; the plan is passed explicitly (no global is
    // touched: other threads may be inside tvc_decoder_infer)
    size_t need = 0;
    for (int impl : {CONV_IMPL_FP32, CONV_IMPL_TC}) {
        Arena A(nullptr, 0, true);
        if (shape_only.infer(A, 0, nullptr, nullptr, nullptr, nullptr, nullptr, B, Lf, impl)) return 0;
        need = A.peak > need ? A.peak : need;
    }
    return need + 256;
}

size_t tvc_decoder_infer_workspace_bytes(int B, int Lf) {
    if (B <= 0 || Lf <= 0) return 0;
    static DecoderModel shape_only;
    Arena A(nullptr, 0, true);
    if (shape_only.infer(A, 0, nullptr, nullptr, nullptr, nullptr, nullptr, B, Lf, g_conv_impl)) return 0;
    return A.peak + 256;
}

int tvc_decoder_infer_range(tvc_decoder_t h, const float* content, const float* f0, const float* energy,
                            const float* rand01, float* out, int B, int Lf, int64_t out_t0, int64_t out_t1, void* workspace,
                            size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && content && f0 && energy && out && workspace, "tvc_decoder_infer: null argument");
    TVC_REQUIRE(Lf <= 0 || (out_t0 >= 0 && out_t0 < out_t1 && out_t1 <= (int64_t)Lf * kFrame),
                "tvc_decoder_infer_range: range [%lld, %lld) outside the %lld output samples", (long long)out_t0, (long long)out_t1,
                (long long)Lf * kFrame);
    const int t0 = (int)out_t0, t1 = (int)out_t1;
    const int impl = g_conv_impl;     // read once: the whole call runs one plan
    TVC_REQUIRE(rand01 || impl == CONV_IMPL_TC, "tvc_decoder_infer: the fp32 plan needs an injected rand01 draw");
    CHECK_SHAPES();
    cudaStream_t s = (cudaStream_t)stream;
    // A caller that is itself capturing this stream (StreamInfer records the whole tick as one graph) gets plain launches:
    // starting / instantiating / launching a graph of our own inside somebody else's capture is not allowed, and the buffer
    // addresses of a capture's private pool can repeat between captures, which would otherwise look like "seen before".
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    TVC_CUDA(cudaStreamIsCapturing(s, &cap_status));
    if (!g_use_graphs || g_prof_on || cap_status != cudaStreamCaptureStatusNone) {
        Arena A(workspace, workspace_bytes, false);
        return h->m.infer(A, s, content, f0, energy, rand01, out, B, Lf, impl, t0, t1);
    }
    // CUDA-graph replay: the launch sequence for one exact set of buffers is captured the second time
    // that set is seen (callers that reuse their buffers -- serving loops, the Python wrapper's cached
    // workspace -- then pay one graph launch instead of ~100 kernel launches per call).
    std::lock_guard<std::mutex> lock(h->mu);
    const DecoderGraphKey key{content, f0, energy, rand01, out, workspace, B, Lf, impl | (g_pdl ? 256 : 0), plan_options(), t0, t1};
    ++h->tick;
    for (DecoderGraph& g : h->graphs)
        if (g.key == key) {
            g.last_use = h->tick;
            TVC_CUDA(cudaGraphLaunch(g.exec, s));
            g_launches.fetch_add(g.launches, std::memory_order_relaxed);
            return 0;
        }
    bool seen_before = false;
    for (const DecoderGraphKey& k : h->seen) seen_before = seen_before || k == key;
    if (!seen_before) {
        if (h->seen.size() >= 32) h->seen.erase(h->seen.begin());
        h->seen.push_back(key);
        Arena A(workspace, workspace_bytes, false);
        return h->m.infer(A, s, content, f0, energy, rand01, out, B, Lf, impl, t0, t1);
    }
    if (!h->cap_stream) TVC_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const unsigned long long n0 = g_launches.load();
    TVC_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int rc = 0;
    {
        Arena A(workspace, workspace_bytes, false);
        rc = h->m.infer(A, h->cap_stream, content, f0, energy, rand01, out, B, Lf, impl, t0, t1);
    }
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
    if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    TVC_REQUIRE(ce == cudaSuccess && graph, "tvc_decoder_infer: graph capture failed: %s", cudaGetErrorString(ce));
    DecoderGraph g;
    g.key = key;
    g.launches = g_launches.load() - n0;
    g.last_use = h->tick;
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    TVC_REQUIRE(ie == cudaSuccess, "tvc_decoder_infer: graph instantiation failed: %s", cudaGetErrorString(ie));
    if (h->graphs.size() >= 8) {               // evict the least recently used
        size_t victim = 0;
        for (size_t i = 1; i < h->graphs.size(); ++i)
            if (h->graphs[i].last_use < h->graphs[victim].last_use) victim = i;
        cudaGraphExecDestroy(h->graphs[victim].exec);
        h->graphs.erase(h->graphs.begin() + victim);
    }
    h->graphs.push_back(g);
    TVC_CUDA(cudaGraphLaunch(g.exec, s));
    return 0;
    API_END
}

int tvc_decoder_infer(tvc_decoder_t h, const float* content, const float* f0, const float* energy,
                      const float* rand01, float* out, int B, int Lf, void* workspace, size_t workspace_bytes,
                      void* stream) {
    return tvc_decoder_infer_range(h, content, f0, energy, rand01, out, B, Lf, 0, (int64_t)(Lf > 0 ? Lf : 1) * kFrame, workspace,
                                   workspace_bytes, stream);
}

int tvc_decoder_plan_windows(int Lf, int64_t out_t0, int64_t out_t1, int32_t* wa, int32_t* wb) {
    API_BEGIN
    TVC_REQUIRE(wa && wb && Lf > 0 && out_t0 >= 0 && out_t0 < out_t1 && out_t1 <= (int64_t)Lf * kFrame, "tvc_decoder_plan_windows: bad arguments");
    int a[5], b[5];
    decoder_plan_windows(Lf, (int)out_t0, (int)out_t1, true, a, b);
    for (int i = 0; i < 5; ++i) { wa[i] = a[i]; wb[i] = b[i]; }
    return 0;
    API_END
}

int tvc_decoder_seed(tvc_decoder_t h, uint64_t seed, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && h->m.tc && h->m.tc->rng_state, "tvc_decoder_seed: decoder has no tensor-core plan");
    return rng_seed(h->m.tc->rng_state, (unsigned long long)seed, (cudaStream_t)stream);
    API_END
}

int tvc_source_net(tvc_decoder_t h, const float* content, const float* f0, const float* energy, float* amps,
                   float* kernel, int B, int Lf, void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && content && f0 && energy && amps && kernel && workspace, "tvc_source_net: null argument");
    CHECK_SHAPES();
    Arena A(workspace, workspace_bytes, false);
    cudaStream_t s = (cudaStream_t)stream;
    float* e_fr = A.f32((int64_t)B * Lf);
    float* lf0 = A.f32((int64_t)B * Lf);
    TVC_REQUIRE(!A.overflow, "workspace too small");
    TVC_TRY(frame_prep(energy, f0, e_fr, lf0, B, Lf, s));
    return h->m.source_net(A, s, content, e_fr, lf0, amps, kernel, B, Lf);
    API_END
}

int tvc_dsp(tvc_decoder_t h, const float* f0, const float* amps, const float* kernel, const float* rand01,
            float* source, int B, int Lf, void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && f0 && amps && kernel && rand01 && source && workspace, "tvc_dsp: null argument");
    CHECK_SHAPES();
    Arena A(workspace, workspace_bytes, false);
    return h->m.dsp(A, (cudaStream_t)stream, f0, amps, kernel, rand01, source, 16LL * Lf * kFrame, B, Lf);
    API_END
}

int tvc_filter_net(tvc_decoder_t h, const float* content, const float* f0, const float* energy, const float* source,
                   float* out, int B, int Lf, void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && content && f0 && energy && source && out && workspace, "tvc_filter_net: null argument");
    CHECK_SHAPES();
    Arena A(workspace, workspace_bytes, false);
    cudaStream_t s = (cudaStream_t)stream;
    const long long L = (long long)Lf * kFrame;
    float* e_fr = A.f32((int64_t)B * Lf);
    float* lf0 = A.f32((int64_t)B * Lf);
    float* src17 = A.f32((int64_t)B * 17 * L);
    TVC_REQUIRE(!A.overflow, "workspace too small");
    TVC_TRY(frame_prep(energy, f0, e_fr, lf0, B, Lf, s));
    TVC_CUDA(cudaMemcpy2DAsync(src17, sizeof(float) * 17 * L, source, sizeof(float) * 16 * L, sizeof(float) * 16 * L, B,
                               cudaMemcpyDeviceToDevice, s));
    TVC_CUDA(cudaMemcpy2DAsync(src17 + 16 * L, sizeof(float) * 17 * L, energy, sizeof(float) * L, sizeof(float) * L, B,
                               cudaMemcpyDeviceToDevice, s));
    return h->m.filter_net(A, s, content, lf0, src17, out, B, Lf);
    API_END
}

int tvc_harmonic_theta(const float* f0, float* theta, int B, int Lf, void* stream) {
    API_BEGIN
    TVC_REQUIRE(f0 && theta, "tvc_harmonic_theta: null argument");
    CHECK_SHAPES();
    return harmonic_theta(f0, theta, B, Lf, (cudaStream_t)stream);
    API_END
}

// ---------------------------------------------------------------------------------------------- encoder
int tvc_encoder_create(const float* params, int64_t numel, tvc_encoder_t* out) {
    API_BEGIN
    TVC_REQUIRE(out, "tvc_encoder_create: null out pointer");
    TVC_TRY(global_init());
    tvc_encoder* h = new (std::nothrow) tvc_encoder();
    TVC_REQUIRE(h, "tvc_encoder_create: out of host memory");
    const int r = h->m.init(params, numel);
    if (r) { delete h; return r; }
    *out = h;
    return 0;
    API_END
}
int tvc_encoder_destroy(tvc_encoder_t h) { delete h; return 0; }

size_t tvc_encoder_workspace_bytes(int B, int Lf) {
    if (B <= 0 || Lf <= 0) return 0;
    const size_t f = sizeof(float);
    // exact-fp32 plan: x, t1 [B,384,Lf]; t2 [B,768,Lf]; sc [B,768]; logits [B,512,Lf]  (+ alignment slack)
    const size_t fp32_plan = (size_t)B * Lf * f * (384 + 384 + 768 + 512) + (size_t)B * 768 * f + 16 * 256;
    // tensor-core plan: dry run of the same launch sequence
    static const EncoderTC shape_only = [] {
        EncoderTC e;
        e.ssl.C = 384; e.ssl.c0 = 0; e.ssl.mid.resize(6); e.ssl.out.Cout = 768;
        e.pitch.C = 128; e.pitch.c0 = 384; e.pitch.mid.resize(4); e.pitch.out.Cout = 512;
        return e;
    }();
    Arena A(nullptr, 0, true);
    float dummy = 0.f;
    (void)shape_only.forward(A, nullptr, &dummy, &dummy, &dummy, B, Lf);
    // either plan, or the content stack on tensor cores followed by the pitch stack in fp32 (+ the logits buffer)
    const size_t tc_plan = A.peak + (size_t)B * 512 * Lf * f + 16 * 256;
    const size_t pitch_side = (size_t)B * Lf * f * (128 + 128 + 256) + (size_t)B * 256 * f + 16 * 256;
    return (tc_plan > fp32_plan ? tc_plan : fp32_plan) + (size_t)B * 512 * Lf * f + pitch_side;
}

int tvc_encoder_forward(tvc_encoder_t h, const float* spec, float* z, float* logits, float* f0, int B, int Lf,
                        void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && spec && workspace, "tvc_encoder_forward: null argument");
    CHECK_SHAPES();
    Arena A(workspace, workspace_bytes, false);
    cudaStream_t s = (cudaStream_t)stream;
    if (g_encoder_tc) {
        TVC_REQUIRE(h->m.tc && h->m.tc->ready, "tvc_encoder_forward: tensor-core plan not initialised");
        float* lg = logits;
        if (!lg && f0) lg = A.f32((int64_t)B * 512 * Lf);
        TVC_REQUIRE(!A.overflow, "workspace too small");
        if (lg && !g_pitch_tc && z) {
            // Two independent stacks over the same spectrogram: the fp32 pitch stack (CUDA cores, launch-latency bound) goes
            // to a side stream with its own slice of the workspace, the content stack (tensor cores) stays on the caller's.
            if (!h->side) {
                TVC_CUDA(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
                TVC_CUDA(cudaEventCreateWithFlags(&h->fork, cudaEventDisableTiming));
                TVC_CUDA(cudaEventCreateWithFlags(&h->join, cudaEventDisableTiming));
            }
            const size_t pitch_ws = (size_t)B * Lf * sizeof(float) * (128 + 128 + 256) + (size_t)B * 256 * sizeof(float) + 8 * 256;
            void* pw = A.bytes(pitch_ws);
            TVC_REQUIRE(!A.overflow, "workspace too small");
            Arena Ap(pw, pitch_ws, false);
            TVC_CUDA(cudaEventRecord(h->fork, s));
            TVC_CUDA(cudaStreamWaitEvent(h->side, h->fork, 0));
            TVC_TRY(h->m.run_stack(Ap, h->side, h->m.pitch, spec, lg, B, Lf));
            if (f0) TVC_TRY(pitch_decode(lg, f0, B, 512, Lf, h->side));
            TVC_CUDA(cudaEventRecord(h->join, h->side));
            // The content stack leaves some SMs to the pitch stack and gives up programmatic dependent launch (see t_sm_cap in
            // tvc_common.cuh), so that the two chains really run side by side.
            const long long frames = (long long)B * Lf;
            const bool share = g_encoder_share && (g_encoder_share_force || frames >= kEncoderShareMinFrames);
            t_sm_cap = !share ? 0 : g_encoder_share_force ? g_encoder_share_sms : frames <= kEncoderShareFrames ? kEncoderShareSmsShort : kEncoderShareSmsLong;
            t_pdl_suppress = share;
            const int r = h->m.tc->forward(A, s, spec, z, nullptr, B, Lf);
            t_sm_cap = 0;
            t_pdl_suppress = false;
            TVC_CUDA(cudaStreamWaitEvent(s, h->join, 0));       // joined even if the content stack failed to launch
            return r;
        }
        if (z || (lg && g_pitch_tc)) TVC_TRY(h->m.tc->forward(A, s, spec, z, g_pitch_tc ? lg : nullptr, B, Lf));
        if (lg && !g_pitch_tc) TVC_TRY(h->m.run_stack(A, s, h->m.pitch, spec, lg, B, Lf));
        if (f0) TVC_TRY(pitch_decode(lg, f0, B, 512, Lf, s));
        return 0;
    }
    if (z) TVC_TRY(h->m.run_stack(A, s, h->m.ssl, spec, z, B, Lf));
    if (logits || f0) {
        float* lg = logits ? logits : A.f32((int64_t)B * 512 * Lf);
        TVC_REQUIRE(!A.overflow, "workspace too small");
        TVC_TRY(h->m.run_stack(A, s, h->m.pitch, spec, lg, B, Lf));
        if (f0) TVC_TRY(pitch_decode(lg, f0, B, 512, Lf, s));
    }
    return 0;
    API_END
}

int tvc_pitch_decode(const float* logits, float* f0, int B, int Lf, void* stream) {
    API_BEGIN
    TVC_REQUIRE(logits && f0, "tvc_pitch_decode: null argument");
    CHECK_SHAPES();
    return pitch_decode(logits, f0, B, 512, Lf, (cudaStream_t)stream);
    API_END
}

// ---------------------------------------------------------------------------------------------- kNN
constexpr int kScreenMinN = 1024;    // below this the exact CUDA-core product is as fast
constexpr int kScreenNT = 128;       // reference vectors per CTA tile of the screening product

int tvc_index_create(const float* index, int N, int metric, tvc_index_t* out) {
    API_BEGIN
    TVC_REQUIRE(index && out, "tvc_index_create: null argument");
    TVC_REQUIRE(N >= 1, "tvc_index_create: empty index");
    TVC_REQUIRE(metric >= 0 && metric <= 2, "tvc_index_create: unknown metric %d (0 cos, 1 IP, 2 L2)", metric);
    TVC_TRY(global_init());
    tvc_index* h = new (std::nothrow) tvc_index();
    TVC_REQUIRE(h, "tvc_index_create: out of host memory");
    IndexModel& m = h->m;
    m.N = N; m.NP = (int)align_up(N, 4); m.metric = metric;
    int r = 0;
    do {
        if (cudaMalloc(&m.index_w, sizeof(float) * (size_t)kContent * m.NP) != cudaSuccess ||
            cudaMalloc(&m.index_nc, sizeof(float) * (size_t)kContent * N) != cudaSuccess ||
            cudaMalloc(&m.bias, sizeof(float) * (size_t)m.NP) != cudaSuccess) {
            set_error("tvc_index_create: cudaMalloc failed for N=%d", N);
            r = 1;
            break;
        }
        if (cudaMemset(m.index_w, 0, sizeof(float) * (size_t)kContent * m.NP) != cudaSuccess) { r = 1; break; }
        r = knn_prepare(index, m.index_w, m.index_nc, m.bias, kContent, N, m.NP, metric, 0);
        if (!r && cudaStreamSynchronize(0) != cudaSuccess) { set_error("tvc_index_create: prepare kernel failed"); r = 1; }
    } while (0);
    if (r) { delete h; return r; }
    m.as_conv.w = m.index_w; m.as_conv.b = m.bias; m.as_conv.Cin = kContent; m.as_conv.Cout = N; m.as_conv.CoutP = m.NP;
    m.as_conv.K = 1;
    static const bool no_screen = getenv("TVC_KNN_EXACT") != nullptr;      // developer switch: CUDA-core similarity product
    if (metric == 0 && N >= kScreenMinN && !no_screen) {
        // normalised index rows on the host -> tensor-core weight image (Cout = N) and the row-major device copy
        std::vector<float> wcn((size_t)kContent * m.NP), wn((size_t)N * kContent);
        r = cudaMemcpy(wcn.data(), m.index_w, sizeof(float) * wcn.size(), cudaMemcpyDeviceToHost) != cudaSuccess;
        if (!r) {
            for (int c = 0; c < kContent; ++c)
                for (int n = 0; n < N; ++n) wn[(size_t)n * kContent + c] = wcn[(size_t)c * m.NP + n];
            r = tc_pack_conv(wn.data(), nullptr, N, kContent, 1, nullptr, nullptr, 0, 0, kScreenNT, m.tc);
        }
        if (!r) r = cudaMalloc(&m.index_wn, sizeof(float) * wn.size()) != cudaSuccess;
        if (!r) r = cudaMemcpy(m.index_wn, wn.data(), sizeof(float) * wn.size(), cudaMemcpyHostToDevice) != cudaSuccess;
        if (r) { set_error("tvc_index_create: building the tensor-core image failed for N=%d", N); delete h; return 1; }
        m.screened = true;
    }
    *out = h;
    return 0;
    API_END
}
int tvc_index_destroy(tvc_index_t h) { delete h; return 0; }

static int match_chunk_utts(int N, int Lf) {
    const long long budget = 512LL << 20;   // similarity scratch per pass
    long long per_utt = (long long)N * Lf * (long long)sizeof(float);
    long long c = budget / (per_utt > 0 ? per_utt : 1);
    return (int)(c < 1 ? 1 : c);
}

size_t tvc_match_workspace_bytes(tvc_index_t h, int B, int Lf) {
    if (!h || B <= 0 || Lf <= 0) return 0;
    const int N = h->m.N;
    const size_t f = sizeof(float);
    const size_t rows = (size_t)B * Lf;
    size_t tot = 0;
    tot += align_up(rows * kContent * f, 256);                               // normalised queries
    tot += align_up(rows * 8 * sizeof(int), 256);                            // indices
    // exact path (IP / L2 / k > 4 / small N): similarity chunk + partial top-k lists
    const int bc = std::min(B, match_chunk_utts(N, Lf));
    size_t exact = align_up((size_t)bc * N * Lf * f, 256) + 2 * align_up((size_t)16 * bc * Lf * 8 * f, 256);
    // screened path: query planes, candidates, flags (the similarity matrix stays on chip)
    // (a split sweep keeps up to 8 candidate lists + scores per query, and only for fewer than 148 x 128 queries)
    const size_t lists = std::min(rows, (size_t)148 * 128) * 8 * 8;
    size_t screened = 2 * align_up(rows * kContent * sizeof(bf16), 256) + align_up(rows * 8 * sizeof(int), 256) +
                      align_up(rows * sizeof(int), 256) + 2 * align_up(lists * sizeof(int), 256);
    tot += h->m.screened ? std::max(exact, screened) : exact;
    return tot + 1024;
}

int tvc_match_features(tvc_index_t h, const float* source, float* out, int32_t* idx_out, int B, int Lf, int k,
                       float alpha, void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(h && source && out && workspace, "tvc_match_features: null argument");
    CHECK_SHAPES();
    const IndexModel& m = h->m;
    TVC_REQUIRE(k >= 1 && k <= 8, "match_features: k=%d unsupported (1..8)", k);
    TVC_REQUIRE(k <= m.N, "match_features: k=%d exceeds the index size %d", k, m.N);
    cudaStream_t s = (cudaStream_t)stream;
    Arena A(workspace, workspace_bytes, false);
    const long long rows = (long long)B * Lf;
    float* qn = A.f32(rows * kContent);
    int* idx = idx_out ? idx_out : A.i32(rows * k);
    TVC_REQUIRE(!A.overflow, "workspace too small: need at least %zu bytes, got %zu", A.peak, A.cap);
    TVC_TRY(knn_normalize_queries(source, qn, B, kContent, Lf, m.metric, s));
    if (m.screened && k <= 4) {
        // Similarity product on the tensor cores with the top candidates selected in its epilogue (the CTA that owns 128
        // queries sweeps the whole index and keeps 8 candidates per query in registers), then exact fp32 re-scoring of
        // the nominated candidates (knn.cu).  Nothing of size queries x N is ever written.
        bf16* q_hi = (bf16*)A.bytes((size_t)rows * kContent * sizeof(bf16));
        bf16* q_lo = (bf16*)A.bytes((size_t)rows * kContent * sizeof(bf16));
        const int splits = tc_conv_topk_splits(m.tc, rows);      // few queries (a streaming tick): cut the sweep so it covers the SMs
        int* cand = A.i32(rows * 8 * splits);
        float* cscore = splits > 1 ? A.f32(rows * 8 * splits) : nullptr;
        int* flag = A.i32(rows);
        TVC_REQUIRE(!A.overflow, "workspace too small: need at least %zu bytes, got %zu", A.peak, A.cap);
        TVC_TRY(cf_to_planes(qn, q_hi, q_lo, B, kContent, Lf, kContent, TC_ACT_NONE, s));
        TcConvArgs a;
        a.a_hi = q_hi; a.a_lo = q_lo; a.a_cs = kContent; a.B = B; a.T = Lf;
        a.topk_cand = cand; a.topk_flag = flag; a.topk_k = k; a.topk_n = m.N; a.topk_eps = kKnnScreenEps;
        a.topk_score = cscore; a.topk_splits = splits;
        {
            ProfScope ps("tc_knn_screen(", s);
            TVC_TRY(tc_conv_launch(m.tc, a, s));
        }
        TVC_TRY(knn_rescore_candidates(qn, m.index_wn, cand, cscore, splits, flag, idx, B, Lf, m.N, k, s));
        return knn_gather_mean(source, m.index_nc, idx, out, B, kContent, Lf, k, alpha, s);
    }
    const int bc = std::min(B, match_chunk_utts(m.N, Lf));
    float* sims = A.f32((int64_t)bc * (int64_t)m.N * Lf);
    float* pv = A.f32((int64_t)16 * bc * Lf * 8);
    int* pi = A.i32((int64_t)16 * bc * Lf * 8);
    TVC_REQUIRE(!A.overflow, "workspace too small: need at least %zu bytes, got %zu", A.peak, A.cap);
    for (int b0 = 0; b0 < B; b0 += bc) {
        const int nb = std::min(bc, B - b0);
        TVC_TRY(conv_run(A, s, m.as_conv, qn + (long long)b0 * kContent * Lf, (long long)kContent * Lf, sims,
                         (long long)m.N * Lf, nb, Lf, 1, PRE_NONE, EPI_NONE));
        TVC_TRY(knn_topk(sims, pv, pi, idx + (long long)b0 * Lf * k, nb, Lf, m.N, k, s));
    }
    return knn_gather_mean(source, m.index_nc, idx, out, B, kContent, Lf, k, alpha, s);
    API_END
}

// ---------------------------------------------------------------------------------------------- front end
size_t tvc_spectrogram_workspace_bytes(int B, int L) {
    if (B <= 0 || L <= 0 || L % kFrame) return 0;
    Arena A(nullptr, 0, true);
    if (spectrogram_run(A, 0, nullptr, nullptr, B, L)) return 0;
    return A.peak + 256;
}
int tvc_spectrogram(const float* wf, float* spec, int B, int L, void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(wf && spec && workspace, "tvc_spectrogram: null argument");
    TVC_REQUIRE(B > 0 && L > 0, "tvc_spectrogram: invalid shape B=%d L=%d", B, L);
    TVC_TRY(global_init());
    Arena A(workspace, workspace_bytes, false);
    return spectrogram_run(A, (cudaStream_t)stream, wf, spec, B, L);
    API_END
}

size_t tvc_energy_workspace_bytes(int B, int L) {
    if (B <= 0 || L < 64) return 0;
    return align_up((size_t)B * energy_pooled_len(L) * sizeof(float), 256) + 256;
}
int tvc_estimate_energy(const float* wf, float* energy, int B, int L, void* workspace, size_t workspace_bytes,
                        void* stream) {
    API_BEGIN
    TVC_REQUIRE(wf && energy && workspace, "tvc_estimate_energy: null argument");
    TVC_REQUIRE(B > 0 && L > 0, "tvc_estimate_energy: invalid shape B=%d L=%d", B, L);
    Arena A(workspace, workspace_bytes, false);
    float* pooled = A.f32((int64_t)B * energy_pooled_len(L));
    TVC_REQUIRE(!A.overflow, "workspace too small: need at least %zu bytes, got %zu", A.peak, A.cap);
    return energy_estimate(wf, energy, pooled, B, L, (cudaStream_t)stream);
    API_END
}

int64_t tvc_resample_length(int64_t L, int orig_freq, int new_freq) {
    if (L < 0 || orig_freq <= 0 || new_freq <= 0) return -1;
    return (int64_t)resample_length((long long)L, orig_freq, new_freq);
}
int tvc_resample(const float* wf, float* out, int B, int64_t L, int orig_freq, int new_freq, void* stream) {
    API_BEGIN
    TVC_REQUIRE(wf && out, "tvc_resample: null argument");
    TVC_REQUIRE(B > 0 && L > 0, "tvc_resample: invalid shape B=%d L=%lld", B, (long long)L);
    return resample_run(wf, out, B, (long long)L, orig_freq, new_freq, (cudaStream_t)stream);
    API_END
}

int tvc_shift_frequency(const float* f0, float* out, int64_t n, float semitones, void* stream) {
    API_BEGIN
    TVC_REQUIRE(f0 && out, "tvc_shift_frequency: null argument");
    return shift_frequency(f0, out, n, semitones, (cudaStream_t)stream);
    API_END
}

// ---------------------------------------------------------------------------------------------- streaming
int tvc_sola(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block, int32_t* shift_out,
             int S, int block, int cross, int search, int delay, void* stream) {
    API_BEGIN
    TVC_REQUIRE(y && sola_buf && fade_in && out_block && shift_out, "tvc_sola: null argument");
    TVC_REQUIRE(S > 0 && block > 0 && cross > 0 && search >= 0 && delay >= 0, "tvc_sola: invalid sizes");
    return sola_run(y, y_len, sola_buf, fade_in, out_block, shift_out, S, block, cross, search, delay, (cudaStream_t)stream);
    API_END
}

size_t tvc_sola_pv_workspace_bytes(int S, int cross) {
    if (S <= 0 || cross <= 0) return 0;
    return sizeof(float) * ((size_t)S * 2 * cross + phase_vocoder_scratch_floats(S, cross)) + 256;
}

int tvc_sola_pv(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block, int32_t* shift_out,
                int S, int block, int cross, int search, int delay, void* workspace, size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(y && sola_buf && fade_in && out_block && shift_out && workspace, "tvc_sola_pv: null argument");
    TVC_REQUIRE(S > 0 && block > 0 && cross > 0 && search >= 0 && delay >= 0, "tvc_sola_pv: invalid sizes");
    TVC_REQUIRE(workspace_bytes >= tvc_sola_pv_workspace_bytes(S, cross), "tvc_sola_pv: workspace too small");
    return sola_run(y, y_len, sola_buf, fade_in, out_block, shift_out, S, block, cross, search, delay, (cudaStream_t)stream,
                    (float*)workspace);
    API_END
}

size_t tvc_phase_vocoder_workspace_bytes(int S, int n) {
    if (S <= 0 || n <= 0) return 0;
    return sizeof(float) * phase_vocoder_scratch_floats(S, n) + 256;
}

int tvc_phase_vocoder(const float* a, const float* b, const float* fade_in, float* out, int S, int n, void* workspace,
                      size_t workspace_bytes, void* stream) {
    API_BEGIN
    TVC_REQUIRE(a && b && fade_in && out && workspace, "tvc_phase_vocoder: null argument");
    TVC_REQUIRE(S > 0 && n > 0, "tvc_phase_vocoder: invalid sizes");
    TVC_REQUIRE(workspace_bytes >= tvc_phase_vocoder_workspace_bytes(S, n), "tvc_phase_vocoder: workspace too small");
    // a and b are separate [S][n] arrays: address b relative to a
    return phase_vocoder_run(a, (long long)n, (long long)(b - a), fade_in, (float*)workspace, out, (long long)n, S, n,
                             (cudaStream_t)stream);
    API_END
}

// ---------------------------------------------------------------------------------------------- parity probes
int tvc_tc_conv_probe(const float* x, const float* w, const float* bias, int B, int T, int Cin, int Cout, int K, int dil,
                      const float* aux_x, const float* aux_w, const float* aux_b, int aux_cin, int aux_mode,
                      const float* res, int epi_act, int out_act, int NT, float* y, float* y_planes, void* stream) {
    API_BEGIN
    TVC_TRY(global_init());
    TVC_REQUIRE(x && w && B > 0 && T > 0 && Cin > 0 && Cout > 0, "tvc_tc_conv_probe: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    TcConvW W;
    TVC_TRY(tc_pack_conv(w, bias, Cout, Cin, K, aux_w, aux_b, aux_cin, aux_mode, NT, W));
    const long long rows = (long long)B * T;
    const int a_cs = (int)align_up(Cin, 8), x_cs = (int)align_up(aux_cin > 0 ? aux_cin : 8, 8), o_cs = (int)align_up(Cout, 8);
    bf16 *a_hi = nullptr, *a_lo = nullptr, *x_hi = nullptr, *x_lo = nullptr, *y_hi = nullptr, *y_lo = nullptr;
    float *res_cl = nullptr, *y_cl = nullptr;
    int rc = 0;
    auto cleanup = [&] {
        cudaStreamSynchronize(s);
        cudaFree(a_hi); cudaFree(a_lo); cudaFree(x_hi); cudaFree(x_lo); cudaFree(y_hi); cudaFree(y_lo);
        cudaFree(res_cl); cudaFree(y_cl);
        W.free_all();
    };
#define PROBE_CUDA(e) do { if ((e) != cudaSuccess) { set_error("probe: %s failed: %s", #e, cudaGetErrorString(cudaGetLastError())); cleanup(); return 1; } } while (0)
#define PROBE_TRY(e) do { rc = (e); if (rc) { cleanup(); return rc; } } while (0)
    PROBE_CUDA(cudaMalloc(&a_hi, rows * a_cs * 2)); PROBE_CUDA(cudaMalloc(&a_lo, rows * a_cs * 2));
    PROBE_CUDA(cudaMalloc(&y_hi, rows * o_cs * 2)); PROBE_CUDA(cudaMalloc(&y_lo, rows * o_cs * 2));
    PROBE_CUDA(cudaMalloc(&y_cl, rows * o_cs * 4));
    PROBE_CUDA(cudaMemsetAsync(y_cl, 0, rows * o_cs * 4, s));
    PROBE_CUDA(cudaMemsetAsync(y_hi, 0, rows * o_cs * 2, s)); PROBE_CUDA(cudaMemsetAsync(y_lo, 0, rows * o_cs * 2, s));
    PROBE_TRY(cf_to_planes(x, a_hi, a_lo, B, Cin, T, a_cs, TC_ACT_NONE, s));
    TcConvArgs a;
    a.a_hi = a_hi; a.a_lo = a_lo; a.a_cs = a_cs; a.dil = dil; a.B = B; a.T = T;
    if (aux_mode != TC_AUX_NONE) {
        TVC_REQUIRE(aux_x && aux_w, "tvc_tc_conv_probe: aux input missing");
        PROBE_CUDA(cudaMalloc(&x_hi, rows * x_cs * 2)); PROBE_CUDA(cudaMalloc(&x_lo, rows * x_cs * 2));
        PROBE_TRY(cf_to_planes(aux_x, x_hi, x_lo, B, aux_cin, T, x_cs, TC_ACT_NONE, s));
        a.x_hi = x_hi; a.x_lo = x_lo; a.x_cs = x_cs;
    }
    if (res) {
        PROBE_CUDA(cudaMalloc(&res_cl, rows * o_cs * 4));
        PROBE_TRY(cf_to_cl(res, res_cl, B, Cout, T, o_cs, s));
        a.res = res_cl; a.res_cs = o_cs;
    }
    a.y32 = y_cl; a.y32_cs = o_cs; a.y_hi = y_hi; a.y_lo = y_lo; a.y_cs = o_cs;
    a.epi_act = epi_act; a.out_act = out_act;
    // tvc_set_option("probe_pad", "Pin,Pout"): run the conv in padded mode (tc_conv.cuh) -- the input planes are re-laid
    // with Pin stored replicate rows, the plane output is produced with Pout and checked: stripping and re-padding it
    // must reproduce it bit for bit (i.e. the epilogue wrote every replicate row), then it is stripped for the caller.
    bf16 *ap_hi = nullptr, *ap_lo = nullptr, *yp_hi = nullptr, *yp_lo = nullptr, *yq = nullptr;
    const int pin = K == 3 ? g_probe_pad_in : 0, pout = K == 3 ? g_probe_pad_out : 0;
    auto cleanup2 = [&] { cudaStreamSynchronize(s); cudaFree(ap_hi); cudaFree(ap_lo); cudaFree(yp_hi); cudaFree(yp_lo); cudaFree(yq); };
    if (pin > 0) {
        const long long rows_pi = (long long)B * (T + 2 * pin), rows_po = (long long)B * (T + 2 * pout);
        if (cudaMalloc(&ap_hi, rows_pi * a_cs * 2) != cudaSuccess || cudaMalloc(&ap_lo, rows_pi * a_cs * 2) != cudaSuccess ||
            cudaMalloc(&yp_hi, rows_po * o_cs * 2) != cudaSuccess || cudaMalloc(&yp_lo, rows_po * o_cs * 2) != cudaSuccess ||
            cudaMalloc(&yq, rows_po * o_cs * 2) != cudaSuccess) {
            set_error("probe: cudaMalloc failed");
            cleanup2(); cleanup();
            return 1;
        }
        cudaMemsetAsync(yp_hi, 0xff, rows_po * o_cs * 2, s);
        cudaMemsetAsync(yp_lo, 0xff, rows_po * o_cs * 2, s);
        rc = plane_repad(a_hi, ap_hi, B, T, a_cs, 0, pin, s);
        if (!rc) rc = plane_repad(a_lo, ap_lo, B, T, a_cs, 0, pin, s);
        a.a_hi = ap_hi; a.a_lo = ap_lo; a.a_pad = pin;
        a.y_hi = yp_hi; a.y_lo = yp_lo; a.y_pad = pout;
        if (!rc) rc = tc_conv_launch(W, a, s);
        // strip -> y_hi / y_lo (what the caller reads), re-pad -> compare with what the kernel wrote
        for (int pl = 0; pl < 2 && !rc; ++pl) {
            bf16* padded = pl ? yp_lo : yp_hi;
            bf16* plain = pl ? y_lo : y_hi;
            rc = plane_repad(padded, plain, B, T, o_cs, pout, 0, s);
            if (!rc) rc = plane_repad(plain, yq, B, T, o_cs, 0, pout, s);
            if (!rc) {
                const size_t n = (size_t)rows_po * o_cs;
                const size_t nc = (size_t)rows_po * (size_t)(align_up(Cout, 8));      // chunks past Cout are never written
                std::vector<uint16_t> h1(n), h2(n);
                cudaStreamSynchronize(s);
                cudaMemcpy(h1.data(), padded, n * 2, cudaMemcpyDeviceToHost);
                cudaMemcpy(h2.data(), yq, n * 2, cudaMemcpyDeviceToHost);
                if (memcmp(h1.data(), h2.data(), nc * 2) != 0) {
                    size_t bad = 0, firstbad = n;
                    for (size_t i = 0; i < nc; ++i)
                        if (h1[i] != h2[i]) { if (firstbad == n) firstbad = i; ++bad; }
                    set_error("probe: padded plane output: %zu elements differ from the replicate-padded reference (first at %zu, plane %d)",
                              bad, firstbad, pl);
                    rc = 1;
                }
            }
        }
        if (rc) { cleanup2(); cleanup(); return rc; }
    } else {
        PROBE_TRY(tc_conv_launch(W, a, s));
    }
    if (y) PROBE_TRY(cl_to_cf(y_cl, y, B, Cout, T, o_cs, s));
    if (y_planes) PROBE_TRY(planes_to_cf(y_hi, y_lo, y_planes, B, Cout, T, o_cs, s));
    PROBE_CUDA(cudaStreamSynchronize(s));
    cleanup2();
    cleanup();
    return 0;
#undef PROBE_CUDA
#undef PROBE_TRY
    API_END
}

}  // extern "C"
#pragma GCC visibility pop

// Shared declarations for the TinyVC sm_100a kernels (internal; the public C-ABI is include/tinyvc_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <mutex>

namespace tvc {

// Thread-local error string returned by tvc_last_error().
void set_error(const char* fmt, ...);

#define TVC_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::tvc::set_error("%s -> %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)
// Every kernel launch of this library goes through TVC_LAUNCH_CHECK(): it counts the launch
// (tvc_launch_count(), reported by bench.py as `gpu_launches`) and surfaces launch errors.
void count_launch();
#define TVC_LAUNCH_CHECK()            \
    do {                              \
        ::tvc::count_launch();        \
        TVC_CUDA(cudaGetLastError()); \
    } while (0)

// Optional per-kernel CUDA-event timing (tvc_set_option("profile","1"); tvc_profile_report()).
// A ProfScope brackets one launcher call with two events on the launch stream; disabled = no-op.
struct ProfScope {
    int slot = -1;
    bool nvtx = false;             // tvc_set_option("nvtx","1"): the scope is also an NVTX range named after the launcher
    cudaStream_t stream = 0;
    ProfScope(const char* name, cudaStream_t s);
    ~ProfScope();
};
#define TVC_TRY(expr)                                                                          \
    do {                                                                                       \
        int r__ = (expr);                                                                      \
        if (r__) return r__;                                                                   \
    } while (0)
#define TVC_REQUIRE(cond, ...)                                                                 \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            ::tvc::set_error(__VA_ARGS__);                                                     \
            return 2;                                                                          \
        }                                                                                      \
    } while (0)

// Runs `f` once per CUDA device (kernel attributes such as the opt-in shared-memory size are per device, so a second
// GPU used by the same process needs its own set-up); f returns 0 on success.
struct PerDeviceOnce {
    std::mutex mu;
    unsigned long long done = 0;
    template <class F>
    int run(F&& f) {
        int dev = 0;
        TVC_CUDA(cudaGetDevice(&dev));
        TVC_REQUIRE(dev >= 0 && dev < 64, "unsupported device ordinal %d", dev);
        std::lock_guard<std::mutex> lock(mu);
        if ((done >> dev) & 1ull) return 0;
        const int r = f();
        if (!r) done |= 1ull << dev;
        return r;
    }
};

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): kernels of the decoder plan are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the next kernel's CTAs may become resident and run
// their prologue (barrier init, TMEM allocation, weight prefetch) while the previous kernel drains.
// Every such kernel must execute griddepcontrol.wait before its first access to memory the previous kernels
// produce or still read (the wait returns once all prerequisite grids have completed and flushed).
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#define TVC_PDL_PROLOGUE()                  \
    do {                                    \
        ::tvc::pdl_launch_dependents();     \
        ::tvc::pdl_wait();                  \
    } while (0)
extern bool g_pdl;      // tvc_set_option("pdl", "0"|"1")
// Per-thread launch shaping while two plans share the GPU (tvc_encoder_forward: the fp32 pitch stack beside the tensor-core
// content stack).  t_sm_cap > 0 caps the grid of the persistent tensor-core kernels, and t_pdl_suppress launches them without
// the programmatic-dependent-launch attribute: with it the NEXT kernel's CTAs take every SM the running one leaves free (they
// wait there for their turn), so the other stream's kernels would find no SM at all.
extern thread_local int t_sm_cap;
extern thread_local bool t_pdl_suppress;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (g_pdl && !t_pdl_suppress) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define TVC_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
    (void)::tvc::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)
#endif

constexpr int kFrame = 480;        // samples per frame @ 24 kHz (decoder.py:242)
constexpr int kNfft = 1920;
constexpr int kBins = 961;         // n_fft/2 + 1
constexpr int kContent = 768;
constexpr int kOsc = 15;           // fundamental + 14 harmonics (decoder.py:243)
constexpr float kSampleRate = 24000.0f;

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// F.interpolate(mode='linear', align_corners=False) source coordinates exactly as ATen's CPU
// kernel evaluates them in fp32 (UpSample.h area_pixel_compute_source_index):
//     src = fma(scale, dst + 0.5, -0.5), clamped at 0;  i0 = floor(src);  l1 = src - i0;  l0 = 1 - l1
// The intrinsics below are never contracted or re-associated by nvcc.
// ---------------------------------------------------------------------------------------------
struct LinCoord {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ LinCoord lin_coord(int dst, float scale, int in_len) {
    float src = __fmaf_rn(scale, __fadd_rn((float)dst, 0.5f), -0.5f);
    src = src < 0.f ? 0.f : src;
    int i0 = (int)floorf(src);
    i0 = i0 < in_len - 1 ? i0 : in_len - 1;
    LinCoord c;
    c.i0 = i0;
    c.i1 = i0 + (i0 < in_len - 1 ? 1 : 0);
    c.l1 = __fsub_rn(src, (float)i0);
    c.l0 = __fsub_rn(1.0f, c.l1);
    return c;
}
// out = fma(x0, l0, rn(x1 * l1))
__device__ __forceinline__ float lin_blend(float x0, float x1, const LinCoord& c) {
    return __fmaf_rn(x0, c.l0, __fmul_rn(x1, c.l1));
}

__device__ __forceinline__ float leaky01(float v) { return v > 0.f ? v : 0.1f * v; }
__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float elu_plus1(float v) { return (v > 0.f ? v : expm1f(v)) + 1.0f; }

// ---------------------------------------------------------------------------------------------
// Generic dense Conv1d (any kernel size 1/3, dilation, replicate padding) -- conv1d.cu
// ---------------------------------------------------------------------------------------------
enum PreOp { PRE_NONE = 0, PRE_LRELU = 1, PRE_AFFINE = 2 };
enum EpiOp { EPI_NONE = 0, EPI_RES = 1, EPI_FILM_RES = 2, EPI_GELU = 3, EPI_ELU1 = 4 };

struct ConvParams {
    const float* x = nullptr;      // [B][Cin][T]   (batch stride x_bs elements)
    long long x_bs = 0;
    const float* w = nullptr;      // packed [K][Cin][CoutP]
    const float* bias = nullptr;   // [Cout] or null
    float* y = nullptr;            // [B][Cout][T]
    long long y_bs = 0;
    const float* res = nullptr;    // EPI_RES / EPI_FILM_RES: [B][Cout][T]
    long long res_bs = 0;
    const float* film = nullptr;   // EPI_FILM_RES: [B][2*Cout][T], rows [0,Cout) scale, [Cout,2Cout) shift
    long long film_bs = 0;
    const float* pre_scale = nullptr;  // PRE_AFFINE: [B][Cin]
    const float* pre_shift = nullptr;  // PRE_AFFINE: [Cin]
    int B = 0, T = 0, Cin = 0, Cout = 0, CoutP = 0, K = 1, dil = 1;
    int pre = PRE_NONE, epi = EPI_NONE;
};
// Launches the conv on `stream`; returns 0 or an error code (message via set_error).
int conv1d_launch(const ConvParams& p, cudaStream_t stream);
int conv1d_init();   // one-time cudaFuncSetAttribute calls

// Conv implementation selector (set through tvc_set_option("conv_impl", ...)).
enum ConvImpl { CONV_IMPL_FP32 = 0, CONV_IMPL_TC = 1 };
extern int g_conv_impl;

}  // namespace tvc

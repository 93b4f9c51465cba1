// Host-side network drivers: parameter tables (torch state_dict order), weight packing, workspace
// arena and the launch sequences for SourceNet / dsp / FilterNet / Encoder.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "tvc_kernels.cuh"

namespace tvc {

struct ParamSpec {
    std::string name;
    int64_t numel = 0;
    int64_t offset = 0;   // element offset into the flat state_dict-ordered buffer
    int d0 = 0, d1 = 0, d2 = 0;
};

struct ParamTable {
    std::vector<ParamSpec> specs;
    std::unordered_map<std::string, int> by_name;
    int64_t total = 0;
    void add(const std::string& name, int d0, int d1 = 0, int d2 = 0);
    void conv(const std::string& prefix, int cout, int cin, int k);   // .weight [cout][cin][k], .bias [cout]
    void convnext(const std::string& prefix, int c);                  // convnext.py:38-47 registration order
    const ParamSpec* find(const std::string& name) const;
};

void build_decoder_table(ParamTable& t);   // decoder.py registration order (Appendix B of SURVEY.md)
void build_encoder_table(ParamTable& t);   // encoder.py registration order

struct ConvW {
    const float* w = nullptr;   // packed [K][Cin][CoutP]
    const float* b = nullptr;   // [Cout]
    int Cin = 0, Cout = 0, CoutP = 0, K = 1;
};

struct CnxtW {
    const float *dw_w = nullptr, *dw_b = nullptr, *ln_g = nullptr, *ln_b = nullptr;
    const float *grn_gamma = nullptr, *grn_beta = nullptr;
    ConvW c2, c3;
    int C = 0, dil = 1;
};

struct WeightStore {
    ParamTable table;
    float* flat = nullptr;     // device copy of the caller's parameters (torch layout)
    float* packed = nullptr;   // conv weights repacked for the kernels
    int64_t packed_cap = 0, packed_used = 0;
    ~WeightStore();
    int load(const float* params, int64_t numel);   // host or device pointer
    const float* raw(const std::string& name) const;
    float* take(int64_t n);                          // carve from `packed` (16-float aligned)
    int make_conv(const std::string& prefix, ConvW& out, cudaStream_t s);
    int make_conv_cat(const std::string& prefix_a, const std::string& prefix_b, ConvW& out, cudaStream_t s);
    int make_cnxt(const std::string& prefix, int C, int dil, CnxtW& out, cudaStream_t s);
};

// Bump allocator over the caller-provided workspace.  In `dry` mode it only measures.
struct Arena {
    char* base = nullptr;
    size_t cap = 0, off = 0, peak = 0;
    bool dry = false, overflow = false;
    Arena(void* p, size_t c, bool d) : base((char*)p), cap(c), dry(d) {}
    float* f32(int64_t n) { return (float*)bytes((size_t)n * sizeof(float)); }
    int* i32(int64_t n) { return (int*)bytes((size_t)n * sizeof(int)); }
    void* bytes(size_t n);
    size_t mark() const { return off; }
    void release(size_t m) { off = m; }
};

struct DecoderTC;   // nets_tc.cuh: tensor-core execution plan

struct DecoderModel {
    WeightStore store;
    DecoderTC* tc = nullptr;
    // SourceNet (decoder.py:102-134)
    ConvW sn_content_in, sn_to_amps, sn_to_kernel;
    const float *sn_energy_w = nullptr, *sn_energy_b = nullptr, *sn_f0_w = nullptr, *sn_f0_b = nullptr;
    CnxtW sn_mid[3];
    // FilterNet (decoder.py:193-233)
    ConvW fn_content_in, fn_down0;
    const float *fn_f0_w = nullptr, *fn_f0_b = nullptr;
    struct Down { ConvW res, c1, c2, c3; int factor = 1; } fn_down[4];
    struct Up { ConvW c1, c2, film1, c3, c4, film2, c5; int factor = 1; } fn_up[5];
    const float *fn_out_w = nullptr, *fn_out_b = nullptr;
    // inverse-rDFT basis for the noise branch
    ConvW dft_cos, dft_sin;
    float* dft_buf = nullptr;
    ~DecoderModel();
    int init(const float* params, int64_t numel);

    int source_net(Arena& A, cudaStream_t s, const float* content, const float* e_fr, const float* lf0, float* amps,
                   float* kern, int B, int Lf);
    int dsp(Arena& A, cudaStream_t s, const float* f0, const float* amps, const float* kern, const float* rand01,
            float* src, long long src_bs, int B, int Lf);
    int filter_net(Arena& A, cudaStream_t s, const float* content, const float* lf0, const float* src17, float* out,
                   int B, int Lf);
    int infer(Arena& A, cudaStream_t s, const float* content, const float* f0, const float* energy,
              const float* rand01, float* out, int B, int Lf, int impl,    // impl: ConvImpl, chosen by the caller
              int out_t0 = 0, int out_t1 = -1);                            // samples that must be produced (tensor-core plan prunes)
};

struct EncoderTC;   // nets_tc.cuh: tensor-core execution plan

struct EncoderModel {
    WeightStore store;
    EncoderTC* tc = nullptr;
    ~EncoderModel();
    struct Stack {
        ConvW in, out;
        const float *ln_g = nullptr, *ln_b = nullptr;
        std::vector<CnxtW> mid;
        int C = 0;
    } ssl, pitch;
    int init(const float* params, int64_t numel);
    int run_stack(Arena& A, cudaStream_t s, const Stack& st, const float* spec, float* out, int B, int Lf);
};

int spectrogram_run(Arena& A, cudaStream_t s, const float* wf, float* spec, int B, int L);   // frontend.cu

int convnext_forward(Arena& A, cudaStream_t s, const CnxtW& L, float* x, float* t1, float* t2, float* sc, int B, int T);
int conv_run(Arena& A, cudaStream_t s, const ConvW& W, const float* x, long long x_bs, float* y, long long y_bs, int B,
             int T, int dil, int pre, int epi, const float* res = nullptr, long long res_bs = 0,
             const float* film = nullptr, long long film_bs = 0, const float* pre_scale = nullptr,
             const float* pre_shift = nullptr);

}  // namespace tvc

// kNN content-feature match (module/tinyvc/feature_retrieval.py:15-33).
//
//   cos : sims = (s / (|s| + 1e-6)) . (r / (|r| + 1e-6))      (:25-27)
//   IP  : sims = s . r                                         (:21)
//   L2  : sims = -cdist(s, r)  -- ranked here by  s.r - |r|^2/2, which orders identically (:23)
//   best = topk(sims, k) ; out = mean_k r[best] * (1 - alpha) + s * alpha      (:28-33)
//
// The similarity matrix of a chunk of utterances is produced by the dense-conv kernel (queries
// are the channels-first activation [B][768][Lf]; the normalised index [768][N] is exactly the
// packed 1x1-conv weight layout with Cout = N), then reduced to the top-k per query by a
// coalesced segment scan + merge, then the un-normalised index rows are gathered and averaged.
#include "nets.cuh"

namespace tvc {

constexpr int kKnnMaxK = 8;
constexpr int kKnnSeg = 16;

// ---- index preparation -----------------------------------------------------------------------
// index_cn: [C][N] (index.pt layout).  Writes index_w [C][NP] (normalised for cos, copy otherwise, zero padded),
// index_nc [N][C] (raw rows for the gather) and bias [N] (L2: -|r|^2/2, else 0).
__global__ void knn_prepare_kernel(const float* __restrict__ index_cn, float* __restrict__ index_w,
                                   float* __restrict__ index_nc, float* __restrict__ bias, int C, int N, int NP,
                                   int metric) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float ss = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = __ldg(index_cn + (long long)c * N + n);
        ss = fmaf(v, v, ss);
        index_nc[(long long)n * C + c] = v;
    }
    const float nrm = __fadd_rn(sqrtf(ss), 1e-6f);
    for (int c = 0; c < C; ++c) {
        const float v = __ldg(index_cn + (long long)c * N + n);
        index_w[(long long)c * NP + n] = metric == 0 ? __fdiv_rn(v, nrm) : v;
    }
    bias[n] = metric == 2 ? -0.5f * ss : 0.f;
}

int knn_prepare(const float* index_cn, float* index_w, float* index_nc, float* bias, int C, int N, int NP, int metric,
                cudaStream_t s) {
    knn_prepare_kernel<<<cdiv(N, 128), 128, 0, s>>>(index_cn, index_w, index_nc, bias, C, N, NP, metric);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---- query normalisation: qn[b,c,t] = s[b,c,t] / (|s[b,:,t]| + 1e-6)   (thread per frame) -----
__global__ void knn_normalize_kernel(const float* __restrict__ src, float* __restrict__ qn, int C, int T, long long ncol,
                                     int metric) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ncol) return;
    const long long b = n / T;
    const int t = (int)(n - b * T);
    const float* sp = src + b * C * (long long)T + t;
    float* qp = qn + b * C * (long long)T + t;
    float ss = 0.f;
    if (metric == 0) {
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(sp + (long long)c * T);
            ss = fmaf(v, v, ss);
        }
    }
    const float nrm = __fadd_rn(sqrtf(ss), 1e-6f);
    for (int c = 0; c < C; ++c) {
        const float v = __ldg(sp + (long long)c * T);
        qp[(long long)c * T] = metric == 0 ? __fdiv_rn(v, nrm) : v;
    }
}

int knn_normalize_queries(const float* src, float* qn, int B, int C, int T, int metric, cudaStream_t s) {
    const long long ncol = (long long)B * T;
    knn_normalize_kernel<<<cdiv(ncol, 128), 128, 0, s>>>(src, qn, C, T, ncol, metric);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---- top-k over the reference axis -------------------------------------------------------------
// sims: [B][N][T].  Pass 1: grid (column tiles, kKnnSeg); thread = one query column, scans its
// segment of references (loads coalesced over t) keeping a sorted top-k.  Ties keep the lower
// reference index.  Pass 2 merges the kKnnSeg partial lists.
__device__ __forceinline__ void topk_insert(float (&v)[kKnnMaxK], int (&id)[kKnnMaxK], int k, float x, int n) {
    if (!(x > v[k - 1])) return;
    int pos = k - 1;
    while (pos > 0 && x > v[pos - 1]) {
        v[pos] = v[pos - 1];
        id[pos] = id[pos - 1];
        --pos;
    }
    v[pos] = x;
    id[pos] = n;
}

__global__ void knn_scan_kernel(const float* __restrict__ sims, float* __restrict__ pv, int* __restrict__ pi, int N,
                                int T, long long ncol, int k) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int seg = blockIdx.y;
    const int per = (N + kKnnSeg - 1) / kKnnSeg;
    const int n0 = seg * per, n1 = min(N, n0 + per);
    const long long b = col / T;
    const int t = (int)(col - b * T);
    const float* sp = sims + b * N * (long long)T + t;
    float v[kKnnMaxK];
    int id[kKnnMaxK];
#pragma unroll
    for (int i = 0; i < kKnnMaxK; ++i) {
        v[i] = -INFINITY;
        id[i] = -1;
    }
    for (int n = n0; n < n1; ++n) topk_insert(v, id, k, __ldg(sp + (long long)n * T), n);
    for (int i = 0; i < k; ++i) {
        pv[((long long)seg * ncol + col) * kKnnMaxK + i] = v[i];
        pi[((long long)seg * ncol + col) * kKnnMaxK + i] = id[i];
    }
}

__global__ void knn_merge_kernel(const float* __restrict__ pv, const int* __restrict__ pi, int* __restrict__ idx_out,
                                 long long ncol, int k) {
    const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    float v[kKnnMaxK];
    int id[kKnnMaxK];
#pragma unroll
    for (int i = 0; i < kKnnMaxK; ++i) {
        v[i] = -INFINITY;
        id[i] = -1;
    }
    // segments are visited in increasing reference order, so strict '>' keeps the lower index on ties
    for (int seg = 0; seg < kKnnSeg; ++seg)
        for (int i = 0; i < k; ++i) {
            const int n = pi[((long long)seg * ncol + col) * kKnnMaxK + i];
            if (n >= 0) topk_insert(v, id, k, pv[((long long)seg * ncol + col) * kKnnMaxK + i], n);
        }
    for (int i = 0; i < k; ++i) idx_out[col * k + i] = id[i];
}

int knn_topk(const float* sims, float* pv, int* pi, int* idx_out, int B, int T, int N, int k, cudaStream_t s) {
    TVC_REQUIRE(k >= 1 && k <= kKnnMaxK, "match_features: k=%d unsupported (1..%d)", k, kKnnMaxK);
    TVC_REQUIRE(k <= N, "match_features: k=%d exceeds the index size %d", k, N);
    const long long ncol = (long long)B * T;
    dim3 grid(cdiv(ncol, 128), kKnnSeg);
    knn_scan_kernel<<<grid, 128, 0, s>>>(sims, pv, pi, N, T, ncol, k);
    TVC_LAUNCH_CHECK();
    knn_merge_kernel<<<cdiv(ncol, 128), 128, 0, s>>>(pv, pi, idx_out, ncol, k);
    TVC_LAUNCH_CHECK();
    return 0;
}

// ---- gather + mean + alpha blend ----------------------------------------------------------------
// out[b,c,t] = (sum_{i<k} index_nc[idx[b,t,i]][c]) / k * (1-alpha) + src[b,c,t] * alpha     (:30-33)
__global__ void knn_gather_kernel(const float* __restrict__ src, const float* __restrict__ index_nc,
                                  const int* __restrict__ idx, float* __restrict__ out, int C, int T, int k,
                                  float one_minus_alpha, float alpha, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int t = (int)(i % T);
    const long long bc = i / T;
    const int c = (int)(bc % C);
    const long long b = bc / C;
    const int* ip = idx + (b * T + t) * k;
    float acc = __ldg(index_nc + (long long)__ldg(ip) * C + c);
    for (int j = 1; j < k; ++j) acc = __fadd_rn(acc, __ldg(index_nc + (long long)__ldg(ip + j) * C + c));
    const float mean = __fdiv_rn(acc, (float)k);
    out[i] = __fadd_rn(__fmul_rn(mean, one_minus_alpha), __fmul_rn(__ldg(src + i), alpha));
}

int knn_gather_mean(const float* src, const float* index_nc, const int* idx, float* out, int B, int C, int T, int k,
                    float alpha, cudaStream_t s) {
    const long long total = (long long)B * C * T;
    knn_gather_kernel<<<cdiv(total, 256), 256, 0, s>>>(src, index_nc, idx, out, C, T, k, (float)(1.0 - (double)alpha),
                                                      alpha, total);
    TVC_LAUNCH_CHECK();
    return 0;
}

}  // namespace tvc

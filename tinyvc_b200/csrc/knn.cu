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

// ---- query normalisation: qn[b,c,t] = s[b,c,t] / (|s[b,:,t]| + 1e-6) ------------------------------
// Block = 32 frames: eight warps stage the [C][32] tile in shared memory (every load of a frame's column in flight at
// once), warp 0 then forms each frame's squared norm as ONE fmaf chain over ascending channels (the order the parity
// fixtures were taken with -- the values now come from shared memory, not from 768 dependent global round trips), and all
// warps divide and store.  A streaming tick (3 584 frames) is 112 blocks.
constexpr int kNormFrames = 32;
__global__ void __launch_bounds__(256) knn_normalize_kernel(const float* __restrict__ src, float* __restrict__ qn, int C, int T,
                                                           long long ncol, int metric) {
    extern __shared__ float tile[];                 // [C][33]
    __shared__ float s_nrm[kNormFrames];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * kNormFrames + lane;
    const bool valid = n < ncol;
    const long long b = valid ? n / T : 0;
    const int t = valid ? (int)(n - b * T) : 0;
    const float* sp = src + b * C * (long long)T + t;
    float* qp = qn + b * C * (long long)T + t;
#pragma unroll 8
    for (int c = warp; c < C; c += 8) tile[c * 33 + lane] = valid ? __ldg(sp + (long long)c * T) : 0.f;
    __syncthreads();
    if (warp == 0) {
        float ss = 0.f;
        if (metric == 0) {
#pragma unroll 8
            for (int c = 0; c < C; ++c) {
                const float v = tile[c * 33 + lane];
                ss = fmaf(v, v, ss);
            }
        }
        s_nrm[lane] = __fadd_rn(sqrtf(ss), 1e-6f);
    }
    __syncthreads();
    if (!valid) return;
    const float nrm = s_nrm[lane];
#pragma unroll 8
    for (int c = warp; c < C; c += 8) {
        const float v = tile[c * 33 + lane];
        qp[(long long)c * T] = metric == 0 ? __fdiv_rn(v, nrm) : v;
    }
}

int knn_normalize_queries(const float* src, float* qn, int B, int C, int T, int metric, cudaStream_t s) {
    const long long ncol = (long long)B * T;
    const size_t smem = sizeof(float) * 33 * (size_t)C;
    TVC_REQUIRE(smem <= 200 * 1024, "knn_normalize: %d channels do not fit the shared-memory tile", C);
    static PerDeviceOnce attr;
    TVC_TRY(attr.run([] { TVC_CUDA(cudaFuncSetAttribute(knn_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); return 0; }));
    knn_normalize_kernel<<<cdiv(ncol, kNormFrames), 256, smem, s>>>(src, qn, C, T, ncol, metric);
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
    // fewer than k comparable similarities (NaN / Inf queries): fall back to row 0 rather than hand the gather index -1
    for (int i = 0; i < k; ++i) idx_out[col * k + i] = id[i] >= 0 ? id[i] : 0;
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


// =============================================================================================
// Tensor-core screening (cos metric, k <= 4).  The similarity product runs on tcgen05 with split-bf16 operands
// (tc_conv.cu; |error| <~ 1e-6 for unit vectors, worst case 3e-5), which is only trusted to nominate candidates:
//   1. approximate sims -> per query the kKnnCand best candidates (ties: lower index), kept in registers by the
//      similarity kernel's epilogue while its CTA sweeps the index (tc_conv.cu, fused top-k): the similarity matrix
//      is never written
//   2. the candidates are re-scored in exact fp32 with the SAME operation order as the CUDA-core path above
//      (acc = fma(w[ci], x[ci], acc), ci ascending), and the top-k of those exact scores is the answer
//   3. if the k-th and the kKnnCand-th approximate scores are closer than kKnnScreenEps (>= 2 x the worst-case
//      screening error) a true top-k member could have been missed; such queries are re-done by an exact full scan.
// With |approx - exact| <= delta and a_(k) - a_(kKnnCand) > 2 delta every exact top-k member is among the
// candidates, so the result is identical to the exact path's.
// =============================================================================================
constexpr int kKnnCand = 8;

__device__ __forceinline__ bool knn_better(float x, int n, float y, int m) { return x > y || (x == y && n < m); }

// exact fp32 score of reference n for query (b, t): the CUDA-core path's operation order (conv1d.cu FMA loop)
__device__ __forceinline__ float knn_exact_score(const float* __restrict__ wrow, const float* __restrict__ q, int T) {
    // one fmaf chain over ascending channels; the loads are independent of it and go out 32 at a time (rows of the
    // normalised index are 16-byte aligned: kContent * 4 bytes per row)
    float acc = 0.f;
    static_assert(kContent % 32 == 0, "batching");
#pragma unroll 1
    for (int c0 = 0; c0 < kContent; c0 += 32) {
        float4 w[8];
        float x[32];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(reinterpret_cast<const float4*>(wrow + c0) + u);
#pragma unroll
        for (int u = 0; u < 32; ++u) x[u] = __ldg(q + (long long)(c0 + u) * T);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc = fmaf(w[u].x, x[4 * u], acc);
            acc = fmaf(w[u].y, x[4 * u + 1], acc);
            acc = fmaf(w[u].z, x[4 * u + 2], acc);
            acc = fmaf(w[u].w, x[4 * u + 3], acc);
        }
    }
    return acc + 0.f;
}

// pass 3: block = 32 queries x kKnnCand candidates.  With a split sweep (S > 1 candidate lists per query, each sorted by
// screening score descending / index ascending) thread (q, 0) first merges them into the query's 8 best -- the list the
// single sweep would have produced -- and decides the "margin too thin" flag from it; then thread (q, j) re-scores
// candidate j exactly and thread (q, 0) selects.
__global__ void __launch_bounds__(32 * kKnnCand) knn_rescore_kernel(const float* __restrict__ qn, const float* __restrict__ index_wn,
                                                                    const int* __restrict__ cand, const float* __restrict__ score,
                                                                    int S, int* __restrict__ flag, float eps, int N,
                                                                    int* __restrict__ idx_out, int T, long long R, int k) {
    __shared__ float sv[32][kKnnCand];
    __shared__ int si[32][kKnnCand];
    const int ql = threadIdx.x / kKnnCand, j = threadIdx.x % kKnnCand;
    const long long row = (long long)blockIdx.x * 32 + ql;
    if (S > 1) {
        if (row < R && j == 0) {
            int head[8] = {0, 0, 0, 0, 0, 0, 0, 0};                      // S <= 8 lists
            const int* cp = cand + row * S * kKnnCand;
            const float* vp = score + row * S * kKnnCand;
            float vk = 0.f, v8 = -INFINITY;
            for (int i = 0; i < kKnnCand; ++i) {
                int w = -1, bn = -1;
                float bx = -INFINITY;
                for (int sp = 0; sp < S; ++sp) {
                    if (head[sp] >= kKnnCand) continue;
                    const int n = cp[sp * kKnnCand + head[sp]];
                    if (n < 0) continue;
                    const float x = vp[sp * kKnnCand + head[sp]];
                    if (w < 0 || x > bx || (x == bx && n < bn)) { w = sp; bn = n; bx = x; }
                }
                if (w >= 0) ++head[w];
                si[ql][i] = bn;
                if (i == k - 1) vk = bx;
                if (i == kKnnCand - 1) v8 = bx;
            }
            flag[row] = (N > kKnnCand && !(vk - v8 > eps)) ? 1 : 0;      // same rule as the single sweep's epilogue
        }
        __syncthreads();
    }
    if (row < R) {
        const int n = S > 1 ? si[ql][j] : cand[row * kKnnCand + j];
        const long long b = row / T;
        const int t = (int)(row - b * T);
        si[ql][j] = n;
        sv[ql][j] = n >= 0 ? knn_exact_score(index_wn + (long long)n * kContent, qn + b * kContent * (long long)T + t, T) : -INFINITY;
    }
    __syncthreads();
    if (row < R && j == 0) {
        // selection sort of the k best by (score desc, index asc) -- torch.topk order on the CPU
        bool used[kKnnCand];
#pragma unroll
        for (int i = 0; i < kKnnCand; ++i) used[i] = si[ql][i] < 0;
        for (int o = 0; o < k; ++o) {
            int best = -1;
            for (int i = 0; i < kKnnCand; ++i)
                if (!used[i] && (best < 0 || knn_better(sv[ql][i], si[ql][i], sv[ql][best], si[ql][best]))) best = i;
            idx_out[row * k + o] = best >= 0 ? si[ql][best] : 0;
            if (best >= 0) used[best] = true;
        }
    }
}

// pass 4 (rare): exact full scan for flagged queries; block = one query
__global__ void __launch_bounds__(256) knn_exact_fallback_kernel(const float* __restrict__ qn, const float* __restrict__ index_wn,
                                                                 const int* __restrict__ flag, int* __restrict__ idx_out, int T,
                                                                 int N, int k) {
    const long long row = blockIdx.x;
    if (!flag[row]) return;
    __shared__ float sv[256][4];
    __shared__ int si[256][4];
    const long long b = row / T;
    const int t = (int)(row - b * T);
    const float* q = qn + b * kContent * (long long)T + t;
    float v[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int id[4] = {-1, -1, -1, -1};
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float x = knn_exact_score(index_wn + (long long)n * kContent, q, T);
        if (id[k - 1] < 0 || knn_better(x, n, v[k - 1], id[k - 1])) {
            int pos = k - 1;
            while (pos > 0 && (id[pos - 1] < 0 || knn_better(x, n, v[pos - 1], id[pos - 1]))) {
                v[pos] = v[pos - 1];
                id[pos] = id[pos - 1];
                --pos;
            }
            v[pos] = x;
            id[pos] = n;
        }
    }
    for (int i = 0; i < 4; ++i) {
        sv[threadIdx.x][i] = v[i];
        si[threadIdx.x][i] = id[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int bi[4] = {-1, -1, -1, -1};
        for (int th = 0; th < 256; ++th)
            for (int i = 0; i < k; ++i) {
                const int n = si[th][i];
                if (n < 0) continue;
                const float x = sv[th][i];
                if (bi[k - 1] < 0 || knn_better(x, n, bv[k - 1], bi[k - 1])) {
                    int pos = k - 1;
                    while (pos > 0 && (bi[pos - 1] < 0 || knn_better(x, n, bv[pos - 1], bi[pos - 1]))) {
                        bv[pos] = bv[pos - 1];
                        bi[pos] = bi[pos - 1];
                        --pos;
                    }
                    bv[pos] = x;
                    bi[pos] = n;
                }
            }
        for (int i = 0; i < k; ++i) idx_out[row * k + i] = bi[i] >= 0 ? bi[i] : 0;
    }
}

int knn_rescore_candidates(const float* qn, const float* index_wn, const int* cand, const float* score, int splits, int* flag,
                           int* idx_out, int B, int T, int N, int k, cudaStream_t s) {
    TVC_REQUIRE(k >= 1 && k <= 4 && k < kKnnCand, "knn_rescore_candidates: k=%d unsupported", k);
    TVC_REQUIRE(splits >= 1 && splits <= 8 && (splits == 1 || score), "knn_rescore_candidates: %d candidate lists per query", splits);
    const long long R = (long long)B * T;
    knn_rescore_kernel<<<cdiv(R, 32), 32 * kKnnCand, 0, s>>>(qn, index_wn, cand, score, splits, flag, kKnnScreenEps, N, idx_out, T, R, k);
    TVC_LAUNCH_CHECK();
    knn_exact_fallback_kernel<<<(unsigned)R, 256, 0, s>>>(qn, index_wn, flag, idx_out, T, N, k);
    TVC_LAUNCH_CHECK();
    return 0;
}

// [C][NP] -> [N][C] (bitwise copy of the normalised index, rows contiguous for the exact re-scoring)
__global__ void knn_transpose_kernel(const float* __restrict__ w, float* __restrict__ wn, int C, int N, int NP) {
    __shared__ float tile[32][33];
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, n = n0 + tx;
        tile[j][tx] = (c < C && n < N) ? __ldg(w + (long long)c * NP + n) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int n = n0 + j, c = c0 + tx;
        if (n < N && c < C) wn[(long long)n * C + c] = tile[tx][j];
    }
}
int knn_transpose_index(const float* index_w, float* index_wn, int C, int N, int NP, cudaStream_t s) {
    knn_transpose_kernel<<<dim3(cdiv(N, 32), cdiv(C, 32)), 256, 0, s>>>(index_w, index_wn, C, N, NP);
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

// PTX wrappers shared by the tcgen05 kernels (tc_conv.cu, tc_block.cu): mbarriers, bulk / tensor TMA copies, cp.async,
// TMEM allocation and access, UMMA descriptors, split-bf16 helpers.  sm_100a only.
#pragma once
#include <cuda.h>      // CUtensorMap (types only)

#include "tc_conv.cuh"

namespace tvc {

// Tensor map of a chunk-major bf16 plane viewed as [nch][ceil(rows / 8)][64 elements]; box = [box_ch][groups][64] (tc_conv.cu)
int tc_make_plane_map(CUtensorMap* m, const bf16* base, long long rows, int nch, int box_ch, int groups);

namespace {

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)      // suspend-time hint (ns): sleep in hardware, do not spin
        : "memory");
    return ok;
}
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA tensor copy global -> shared of one [chunks][G row groups][128 B] box; the box's bytes complete on the mbarrier
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// 16-byte cp.async (LDGSTS) with zero fill when src_bytes == 0
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one (pre-counted) arrival once all prior cp.async of this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// One lane of a converged warp (the MMA warp runs its loops warp-uniformly so that descriptors stay in
// uniform registers; only the tcgen05 instructions themselves are predicated on the elected lane).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, single-CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (stride between the two 8-element
//   K chunks of one MMA) | [32,46) stride byte offset >> 4 (stride between 8-row groups) |
//   [46,48) version = 1 | [61,64) layout type = 0.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format BF16 (1) @7/@10,
// a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24.
__device__ __forceinline__ uint32_t umma_idesc(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case TC_ACT_LRELU: return leaky01(v);
        case TC_ACT_GELU: return gelu_erf(v);
        case TC_ACT_ELU1: return elu_plus1(v);
        default: return v;
    }
}
__device__ __forceinline__ uint32_t pack_bf16x2(bf16 a, bf16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// split two fp32 values into packed bf16 hi / lo pairs
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l = __floats2bfloat162_rn(__fsub_rn(a, ha), __fsub_rn(b, hb));
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// One K-stage of the implicit GEMM: TAPS x KS K-steps, three bf16 products each (hi*hi, hi*lo, lo*hi), straight-line.
// a_lo0 / b_lo0: low descriptor words (start address | LBO << 16, 16-byte units) of the stage's activation window and
// weight image; *_tap / *_ks: start-address increments per tap and per 16-channel K-step; plane_*: hi -> lo plane.
template <int TAPS, int KS>
__device__ __forceinline__ void issue_stage(uint32_t d, uint32_t a_lo0, uint32_t b_lo0, uint32_t a_tap, uint32_t b_tap,
                                            uint32_t a_ks, uint32_t b_ks, uint32_t plane_a16, uint32_t plane_b16,
                                            uint32_t desc_hi, uint32_t idesc, uint32_t acc0) {
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const uint32_t a_lo = a_lo0 + (uint32_t)t * a_tap + (uint32_t)k * a_ks;
            const uint32_t b_lo = b_lo0 + (uint32_t)t * b_tap + (uint32_t)k * b_ks;
            const uint64_t a_h = ((uint64_t)desc_hi << 32) | a_lo, a_l = ((uint64_t)desc_hi << 32) | (a_lo + plane_a16);
            const uint64_t b_h = ((uint64_t)desc_hi << 32) | b_lo, b_l = ((uint64_t)desc_hi << 32) | (b_lo + plane_b16);
            umma_bf16(d, a_h, b_h, idesc, (t == 0 && k == 0) ? acc0 : 1u);
            umma_bf16(d, a_h, b_l, idesc, 1u);
            umma_bf16(d, a_l, b_h, idesc, 1u);
        }
    }
}

// The same K-stage with the weight planes concatenated along N ("cat" images, tc_conv.cuh): per K-step
//   D[:, 0 : 2 NTp) += x_hi * [w_hi | w_lo]        one MMA of N = 2 NTp
//   D[:, 0 : NTp)   += x_lo * w_hi                  one MMA of N = NTp
// i.e. two instructions instead of three; the epilogue adds the two column halves.  A tcgen05.mma costs ~45 cycles of issue
// whatever N <= 64 is (profiles/r02a_umma_bench.log), so small-N layers are issue-bound and gain ~30 %.
// b_lo0's LBO field and b_ks describe the concatenated image (chunk stride 2 NTp rows).
template <int TAPS, int KS>
__device__ __forceinline__ void issue_stage_cat(uint32_t d, uint32_t a_lo0, uint32_t b_lo0, uint32_t a_tap, uint32_t b_tap,
                                                uint32_t a_ks, uint32_t b_ks, uint32_t plane_a16, uint32_t desc_hi,
                                                uint32_t idesc_n, uint32_t idesc_2n, uint32_t acc0) {
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int k = 0; k < KS; ++k) {
            const uint32_t a_lo = a_lo0 + (uint32_t)t * a_tap + (uint32_t)k * a_ks;
            const uint32_t b_lo = b_lo0 + (uint32_t)t * b_tap + (uint32_t)k * b_ks;
            const uint64_t a_h = ((uint64_t)desc_hi << 32) | a_lo, a_l = ((uint64_t)desc_hi << 32) | (a_lo + plane_a16);
            const uint64_t b = ((uint64_t)desc_hi << 32) | b_lo;
            umma_bf16(d, a_h, b, idesc_2n, (t == 0 && k == 0) ? acc0 : 1u);
            umma_bf16(d, a_l, b, idesc_n, 1u);
        }
    }
}

}  // namespace
}  // namespace tvc

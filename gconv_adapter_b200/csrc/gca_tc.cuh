// tcgen05 / TMEM / mbarrier building blocks for the dense pieces of the adapter (sm_100a only).
//
// Why tensor cores at all: ncu on the CUDA-core v1 kernels (profiles/r1_v2_*.csv) shows the dense
// projections at ~33 % FMA-pipe and ~30 % DRAM utilisation, i.e. issue/latency bound; the six skinny
// GEMMs of one fwd+bwd are 12 N d r FLOP = 8.5 GFLOP at arxiv shape, more than the fp32 pipe can
// retire inside the 211 us HBM time of the step.  fp32 parity (rtol 1e-5) forbids plain TF32, so every
// product is split:  x = hi + lo  with hi = tf32(x), lo = tf32(x - hi), and
//     A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo        (relative error ~2^-21 per product)
// as three tcgen05.mma.kind::tf32 instructions accumulating in fp32 TMEM.
//
// Shared-memory operand layout (no swizzle, "interleaved" canonical layout of the UMMA descriptors):
// 8 x 16-byte core matrices.  A [rows x cols] fp32 tile is stored column-chunk-major:
//     addr(row, col) = (col/4) * S_C + (row/8) * 128 + (row%8) * 16 + (col%4) * 4
// with S_C = (rows/8) * 128.  The SAME bytes serve as a K-major operand (M/N = rows, K = cols:
// LBO = S_C, SBO = 128) and as an MN-major operand (M/N = cols, K = rows: LBO = 128, SBO = S_C).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gca {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// Generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads).
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 ----------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version bits.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// 32-bit instruction descriptor: kind::tf32, fp32 accumulate, dense.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive columns (lane field of taddr must be
// 32 * (warp_id % 4)).  Call tmem_ld_wait() before using the values.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tf32 split ------------------------------------------------------------------------------------
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// x = hi + lo (+ <= 2^-22 |x|): hi = x with the 13 low mantissa bits cleared (exactly representable in
// tf32; the truncation error is carried exactly by x - hi), lo = (x - hi) rounded to nearest tf32 so the
// representation error stays unbiased.  One LOP + FADD + IADD + LOP per value (cvt.rna.tf32.f32 compiles to
// a 4-instruction sequence with an inf/nan test).
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float4 x, float4& hi, float4& lo) {
    split1(x.x, hi.x, lo.x); split1(x.y, hi.y, lo.y); split1(x.z, hi.z, lo.z); split1(x.w, hi.w, lo.w);
}

// Byte offset of element (row, col) in a column-chunk-major core-matrix tile with `rows` rows.
__device__ __forceinline__ uint32_t cm_offset(int row, int col, int rows) {
    return (uint32_t)((col >> 2) * ((rows >> 3) * 128) + (row >> 3) * 128 + (row & 7) * 16 + (col & 3) * 4);
}

}  // namespace tc
}  // namespace gca

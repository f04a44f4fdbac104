// K6: deterministic second-stage reduction of all per-CTA partials of the backward + gscalar.
#include "gca_common.cuh"
#include "gca_device.cuh"
#include "gca_host.cuh"

namespace gca {
namespace {

// ------------------------------------------------------------------------------------------
// K6: second-stage reduction of the per-CTA partials, fixed order.  A job = one block of 32 consecutive
// column quads of one partial array (512 contiguous bytes per partial row: coalesced).  The 32 warps of a CTA
// split the partial rows (warp w sums rows w, w+32, ... with 8 loads in flight), their sums are added in warp
// order through shared memory, and warp 0 writes the result.  The last CTA to finish adds up the gscalar pieces.
// ------------------------------------------------------------------------------------------
constexpr int kFinWarps = 32;   // <= 148 + 148 partial rows per array: every warp sums its rows in ONE batch of loads
__device__ __forceinline__ float4 sum_rows_strided(const float* __restrict__ base, int np, size_t pitch, int first, int stride,
                                                   bool ok) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!ok) return acc;
    for (int p = first; p < np; p += 8 * stride) {   // 8 predicated loads in flight, no serial tail
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            v[u] = (p + stride * u < np) ? ldg4(base + (size_t)(p + stride * u) * pitch) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = f4_add(acc, v[u]);
    }
    return acc;
}

__global__ void __launch_bounds__(kFinWarps * 32)
k_finalize(const float* __restrict__ partGu, const float* __restrict__ partCol, const float* __restrict__ partGd,
           const float* __restrict__ partDot, const float* __restrict__ partBd, float* gsp, int* header,
           const float* __restrict__ Wu, const float* __restrict__ bu, const float* __restrict__ scalar, int skip,
           float* gWd, float* gbd, float* gWu, float* gbu, float* gscalar, int d, int r) {
    __shared__ float4 s_part[kFinWarps][32];
    __shared__ float s_red[kFinWarps * 32];
    __shared__ int s_last;
    pdl_wait();
    const int pu = header[0], pd = header[1], pb = header[2];
    const float s = scalar ? __ldg(scalar) : 1.f;
    const int rd4 = (r * d) >> 2, d4 = d >> 2, r4 = r >> 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int jb_g = (rd4 + 31) / 32, jb_c = (d4 + 31) / 32;
    const int njobs = 2 * jb_g + jb_c + 1;             // [gu blocks][gd blocks][colsum blocks][bd]
    float gs = 0.f;
    for (int job = blockIdx.x; job < njobs; job += gridDim.x) {
        int kind, blk;
        if (job < jb_g) { kind = 0; blk = job; }
        else if (job < 2 * jb_g) { kind = 1; blk = job - jb_g; }
        else if (job < 2 * jb_g + jb_c) { kind = 2; blk = job - 2 * jb_g; }
        else { kind = 3; blk = 0; }
        const float* part = kind == 0 ? partGu : kind == 1 ? partGd : kind == 2 ? partCol : partBd;
        const int np = kind == 0 ? pu : kind == 1 ? pd : kind == 2 ? pu : pb;
        const int nq = kind <= 1 ? rd4 : kind == 2 ? d4 : r4;
        const size_t pitch = kind <= 1 ? (size_t)r * d : kind == 2 ? (size_t)d : (size_t)r;
        // the bias partials are only r floats wide but there are many of them (one per CTA of the transpose hop):
        // the 32 / r4 lane groups of a warp take different partial rows and are combined by a fixed butterfly
        const bool narrow = kind == 3 && r4 < 32 && (32 % r4) == 0;
        const int groups = narrow ? 32 / r4 : 1;
        const int q = narrow ? lane % r4 : blk * 32 + lane;
        const int first = narrow ? warp * groups + lane / r4 : warp;
        const bool live = q < nq && !(kind == 1 && !gWd) && !(kind == 3 && !gbd);
        float4 mine = sum_rows_strided(part + (size_t)q * 4, np, pitch, first, kFinWarps * groups, live);
        if (narrow)
            for (int off = r4; off < 32; off <<= 1) mine = f4_add(mine, f4_shfl_xor(mine, off));
        const bool ok = live && (!narrow || lane < r4);
        __syncthreads();
        s_part[warp][lane] = mine;
        __syncthreads();
        if (warp == 0 && ok) {
            float4 t = s_part[0][lane];
#pragma unroll
            for (int w = 1; w < kFinWarps; ++w) t = f4_add(t, s_part[w][lane]);
            if (kind == 0) {
                const int c = (q * 4) / d, k = q * 4 - c * d;
                const float g4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (gWu) gWu[(size_t)(k + j) * r + c] = s * g4[j];
                    gs = fmaf(g4[j], __ldg(Wu + (size_t)(k + j) * r + c), gs);
                }
            } else if (kind == 1) {
                *reinterpret_cast<float4*>(gWd + (size_t)q * 4) = t;
            } else if (kind == 2) {
                if (gbu) *reinterpret_cast<float4*>(gbu + q * 4) = f4_scale(t, s);
                const float4 b = ldg4(bu + q * 4);
                gs = fmaf(t.x, b.x, fmaf(t.y, b.y, fmaf(t.z, b.z, fmaf(t.w, b.w, gs))));
            } else {
                *reinterpret_cast<float4*>(gbd + q * 4) = t;
            }
        }
    }
    if (!gscalar) return;
    auto block_sum = [&](float v) {                    // fixed-order tree over the CTA
        __syncthreads();
        s_red[threadIdx.x] = v;
        __syncthreads();
        for (int o = kFinWarps * 16; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
            __syncthreads();
        }
        return s_red[0];
    };
    if (skip && blockIdx.x == 0)
        for (int p = threadIdx.x; p < pd; p += blockDim.x) gs += partDot[p];
    const float mine_gs = block_sum(gs);
    if (threadIdx.x == 0) {
        gsp[blockIdx.x] = mine_gs;
        __threadfence();
        s_last = (atomicAdd(&header[3], 1) == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {                                      // gridDim.x <= kMaxFin <= blockDim.x
        __threadfence();
        const float v = threadIdx.x < gridDim.x ? reinterpret_cast<volatile float*>(gsp)[threadIdx.x] : 0.f;
        const float total = block_sum(v);
        if (threadIdx.x == 0) {
            *gscalar = total;
            header[3] = 0;
        }
    }
}

}  // namespace

int launch_finalize(const Scratch& S, const float* Wu, const float* bu, const float* scalar, int skip, float* gWd, float* gbd,
                    float* gWu, float* gbu, float* gscalar, int d, int r, cudaStream_t st) {
    const int rd4 = (r * d) / 4;
    int grid = 2 * ((rd4 + 31) / 32) + (d / 4 + 31) / 32 + 1;   // one CTA per job
    if (grid > kMaxFin) grid = kMaxFin;
    {
        ProfScope ps("finalize", st);
        GCA_CUDA(launch_pdl(k_finalize, dim3(grid), dim3(kFinWarps * 32), 0, st, S.gu, S.col, S.gd, S.dot, S.bd,
                            S.gsp, S.header, Wu, bu, scalar, skip, gWd, gbd, gWu, gbu, gscalar, d, r));
    }
    GCA_LAUNCH_OK();
    return GCA_OK;
}

}  // namespace gca

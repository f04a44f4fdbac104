// K4 weight gradient  G[R,d] = H^T A  (+ column sums of A, <A,B>) with per-CTA partials, no atomics; register-fed variants:
//   k_wgrad_stream   3xTF32 mma.sync, one contiguous row range per CTA, H fragments as LDS.128 quads   (r = 16 / 32)
//   k_wgrad          FFMA                                                                               (every other shape)
// The default for large inputs is the TMA-fed family in gca_stream.cu.
// Reference semantics: autograd of the two Linear layers of /root/reference/src/finetune/gconv_adapter.py:92.
#include "gca_common.cuh"
#include "gca_device.cuh"
#include "gca_host.cuh"

namespace gca {
namespace {

// ------------------------------------------------------------------------------------------
// K4: G[c, col0 + k] = sum_i H[i, c] A[i, col0 + k]   (+ colsum[k] = sum_i A[i,k], dot = sum A.B)
// for the dsub columns starting at col0.  Warp = (column chunk of 128, block of CW <= 16 c's, row
// subset); lane = one column quad.  Each CTA keeps its accumulators in registers across all its
// row blocks and emits ONE partial per launch (no atomics, fixed order).
// ------------------------------------------------------------------------------------------
constexpr int kWgRows = 128;   // rows of H staged per block

template <int R>
__global__ void __launch_bounds__(384)
k_wgrad(const float* __restrict__ A, int64_t lda, const float* __restrict__ H, const float* __restrict__ B, int64_t ldb,
        float* __restrict__ partG, float* __restrict__ partCol, float* __restrict__ partDot, int* header, int header_slot,
        int n, int d, int col0, int dsub, int nchunks, int RS) {
    constexpr int CW = R < 16 ? R : 16;   // c's per warp
    constexpr int CB = R / CW;            // c blocks
    extern __shared__ __align__(16) float smem[];
    float* Hs = smem;                     // [kWgRows][R], later reused as G[R][dsub] + colsum[dsub]
    __shared__ float s_dot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int chunk = warp % nchunks;
    const int cb = (warp / nchunks) % CB;
    const int rs = warp / (nchunks * CB);
    const int q = chunk * 32 + lane;      // column quad inside [col0, col0 + dsub)
    const bool active = q < (dsub >> 2);
    const bool lead = cb == 0;            // this warp also owns colsum / dot for its columns
    const float* Ac = A + col0 + q * 4;
    const float* Bc = B ? B + col0 + q * 4 : nullptr;
    float4 acc[CW];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    float dot = 0.f;
    const int nblocks = (n + kWgRows - 1) / kWgRows;
    for (int blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const int row0 = blk * kWgRows;
        const int rows = min(kWgRows, n - row0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < rows * (R / 4); idx += blockDim.x)
            reinterpret_cast<float4*>(Hs)[idx] = ldg4(H + (size_t)row0 * R + idx * 4);
        __syncthreads();
        if (!active) continue;
        for (int i0 = rs; i0 < rows; i0 += 4 * RS) {
            float4 a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * RS;
                if (i < rows) {
                    a[u] = ldg4_stream(Ac + (size_t)(row0 + i) * lda);
                    if (Bc && lead) b[u] = ldg4_stream(Bc + (size_t)(row0 + i) * ldb);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * RS;
                if (i >= rows) break;       // warp-uniform
                const float* hrow = Hs + i * R + cb * CW;
#pragma unroll
                for (int c4 = 0; c4 < CW / 4; ++c4) {
                    const float4 h = *reinterpret_cast<const float4*>(hrow + c4 * 4);
                    const float hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4& g = acc[c4 * 4 + j];
                        g.x = fmaf(hv[j], a[u].x, g.x); g.y = fmaf(hv[j], a[u].y, g.y);
                        g.z = fmaf(hv[j], a[u].z, g.z); g.w = fmaf(hv[j], a[u].w, g.w);
                    }
                }
                if (lead) {
                    csum = f4_add(csum, a[u]);
                    if (Bc) dot = fmaf(a[u].x, b[u].x, fmaf(a[u].y, b[u].y, fmaf(a[u].z, b[u].z, fmaf(a[u].w, b[u].w, dot))));
                }
            }
        }
    }
    // combine the RS row subsets in a fixed order through shared memory
    float* G = smem;                         // [R][dsub]
    float* Cs = smem + (size_t)R * dsub;     // [dsub]
    for (int turn = 0; turn < RS; ++turn) {
        __syncthreads();
        if (rs == turn && active) {
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                float4* g = reinterpret_cast<float4*>(G + (size_t)(cb * CW + c) * dsub + q * 4);
                *g = turn == 0 ? acc[c] : f4_add(*g, acc[c]);
            }
            if (lead) {
                float4* cs = reinterpret_cast<float4*>(Cs + q * 4);
                *cs = turn == 0 ? csum : f4_add(*cs, csum);
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if (lane == 0) s_dot[warp] = dot;
    __syncthreads();
    const int nq = dsub >> 2;
    float* pg = partG + (size_t)blockIdx.x * R * d + col0;
    for (int idx = threadIdx.x; idx < R * nq; idx += blockDim.x) {
        const int c = idx / nq, qq = idx - c * nq;
        *reinterpret_cast<float4*>(pg + (size_t)c * d + qq * 4) = *reinterpret_cast<const float4*>(G + (size_t)c * dsub + qq * 4);
    }
    if (partCol)
        for (int idx = threadIdx.x; idx < nq; idx += blockDim.x)
            *reinterpret_cast<float4*>(partCol + (size_t)blockIdx.x * d + col0 + idx * 4) = *reinterpret_cast<const float4*>(Cs + idx * 4);
    if (threadIdx.x == 0) {
        if (partDot) {
            float t = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_dot[w];
            if (col0 == 0) partDot[blockIdx.x] = t; else partDot[blockIdx.x] += t;   // launches are stream-ordered
        }
        if (blockIdx.x == 0) header[header_slot] = gridDim.x;
    }
}

constexpr int kMmaRows = 128;

// ------------------------------------------------------------------------------------------
// K4-stream: G[c, k] = sum_i H[i, c] A[i, k] (+ colsum, dot) with M = c, N = k (columns of A), K = rows, as one
// uninterrupted stream per CTA: 8 warps = 8 column blocks of 32; every warp sweeps all rows of the CTA for its block.
// The products are accumulated by the tensor core per 128-row chunk, then folded into a running fp32 sum with a plain
// FADD (the tensor core's accumulator truncates; 48 updates per fold keep that bias ~1e-7).  CTA b owns the contiguous rows [b*rows_per, (b+1)*rows_per) (rows_per a multiple of 32: at most one batch
// of imbalance instead of one 128-row tile), walks them in 128-row chunks and keeps the loads of the next batch
// in flight across batch AND chunk boundaries (two register buffers, 8-16 LDG.128 per lane outstanding).  The H
// chunk lives in shared memory as ready-made A-fragment quads {(kr,g),(kr,g+8),(kr+4,g),(kr+4,g+8)} per lane, hi
// and lo (one conflict-free LDS.128 each per k-step and m-tile), double-buffered: chunk c+1 is fetched into
// registers while chunk c is multiplied and stored after it, one __syncthreads per chunk.  The running fp32
// sums sit in shared memory (the tensor-core accumulator is folded into them once per chunk, see above).
// ------------------------------------------------------------------------------------------
template <int R>
constexpr size_t wgrad_stream_smem() { return (size_t)(4 * 16 * (R / 16) * 32 + 4 * (R / 16) * 256) * 16; }

template <int R, bool HAS_B>
__device__ __forceinline__ void wgrad_stream_body(const float* __restrict__ A, int64_t lda, const float* __restrict__ H,
                                                  const float* __restrict__ B, int64_t ldb, float* __restrict__ partG,
                                                  float* __restrict__ partCol, float* __restrict__ partDot, int* header,
                                                  int header_slot, int n, int d, int col_base, int bid, int nblocks,
                                                  uint4* smem) {
    constexpr int MT = R / 16;
    constexpr int BR = HAS_B ? 16 : 32;       // rows per load batch (8 LDG.128 per lane either way)
    constexpr int KS = BR / 8;                // k-steps per batch
    constexpr int NB = kMmaRows / BR;         // batches per chunk
    constexpr int QPB = 16 * MT * 32;         // fragment quads per H buffer
    constexpr int HT = QPB / 256;             // staging tasks per thread
    uint4* Hh = smem;                         // [2][QPB]
    uint4* Hl = smem + 2 * QPB;               // [2][QPB]
    float4* run = reinterpret_cast<float4*>(smem + 4 * QPB);   // [4 * MT][256]
    __shared__ float s_dot[8];
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int cb = col_base + warp * 32;
    const int col = cb + 4 * g;
    const bool col_ok = col < d;
    const int rows_per = (((n + nblocks - 1) / nblocks) + 31) & ~31;
    const int rb = min(n, bid * rows_per), re = min(n, rb + rows_per);
    const int nch = (re - rb + kMmaRows - 1) / kMmaRows;
#pragma unroll
    for (int q = 0; q < 4 * MT; ++q) run[q * 256 + threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);
    float dot = 0.f;

    float hreg[HT][4];
    auto hload = [&](int row0) {              // task idx = (ks * MT + m) * 32 + (4 g_ + t_)
#pragma unroll
        for (int i = 0; i < HT; ++i) {
            const int idx = threadIdx.x + 256 * i;
            const int km = idx >> 5, ks = km / MT, m = km - ks * MT;
            const int r0_ = row0 + ks * 8 + t, r1_ = r0_ + 4, c = m * 16 + g;
            hreg[i][0] = r0_ < re ? __ldg(H + (size_t)r0_ * R + c) : 0.f;
            hreg[i][1] = r0_ < re ? __ldg(H + (size_t)r0_ * R + c + 8) : 0.f;
            hreg[i][2] = r1_ < re ? __ldg(H + (size_t)r1_ * R + c) : 0.f;
            hreg[i][3] = r1_ < re ? __ldg(H + (size_t)r1_ * R + c + 8) : 0.f;
        }
    };
    auto hstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < HT; ++i) {
            uint4 hi, lo;
            split_tf32(hreg[i][0], hi.x, lo.x); split_tf32(hreg[i][1], hi.y, lo.y);
            split_tf32(hreg[i][2], hi.z, lo.z); split_tf32(hreg[i][3], hi.w, lo.w);
            Hh[buf * QPB + threadIdx.x + 256 * i] = hi;
            Hl[buf * QPB + threadIdx.x + 256 * i] = lo;
        }
    };
    float4 va[KS][2], vb[KS][2], wa[KS][2], wb[KS][2];
    auto load = [&](float4 (&v)[KS][2], float4 (&w)[KS][2], int row0) {
#pragma unroll
        for (int u = 0; u < KS; ++u)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int r_ = row0 + u * 8 + t + 4 * h2;
                const bool ok = col_ok && r_ < re;
                v[u][h2] = ok ? ldg4_stream(A + (size_t)r_ * lda + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (HAS_B) w[u][h2] = ok ? ldg4_stream(B + (size_t)r_ * ldb + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
    };
    float acc[MT][4][4];
    auto compute = [&](const float4 (&v)[KS][2], const float4 (&w)[KS][2], int row0, int ks0, const uint4* hh, const uint4* hl) {
#pragma unroll
        for (int u = 0; u < KS; ++u) {
            if (row0 + u * 8 >= re) break;                   // warp-uniform
            uint32_t bh[2][4], bl[2][4];
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                split_tf32(v[u][h2].x, bh[h2][0], bl[h2][0]); split_tf32(v[u][h2].y, bh[h2][1], bl[h2][1]);
                split_tf32(v[u][h2].z, bh[h2][2], bl[h2][2]); split_tf32(v[u][h2].w, bh[h2][3], bl[h2][3]);
                csum = f4_add(csum, v[u][h2]);
                if (HAS_B) dot = fmaf(v[u][h2].x, w[u][h2].x, fmaf(v[u][h2].y, w[u][h2].y,
                                 fmaf(v[u][h2].z, w[u][h2].z, fmaf(v[u][h2].w, w[u][h2].w, dot))));
            }
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const uint4 qh = hh[((ks0 + u) * MT + m) * 32 + lane];
                const uint4 ql = hl[((ks0 + u) * MT + m) * 32 + lane];
                const uint32_t ah[4] = {qh.x, qh.y, qh.z, qh.w}, al[4] = {ql.x, ql.y, ql.z, ql.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mma_tf32(acc[m][j], ah, bh[0][j], bh[1][j]);
                    mma_tf32(acc[m][j], al, bh[0][j], bh[1][j]);
                    mma_tf32(acc[m][j], ah, bl[0][j], bl[1][j]);
                }
            }
        }
    };
    if (nch > 0) {
        hload(rb);
        load(va, wa, rb);
        hstore(0);
    }
    for (int c = 0; c < nch; ++c) {
        const int row_c = rb + c * kMmaRows;
        const uint4* hh = Hh + (c & 1) * QPB;
        const uint4* hl = Hl + (c & 1) * QPB;
        __syncthreads();                                     // chunk c staged; everyone is done with chunk c-1
        if (c + 1 < nch) hload(row_c + kMmaRows);
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[m][j][i] = 0.f;
#pragma unroll
        for (int b = 0; b < NB; b += 2) {
            load(vb, wb, row_c + (b + 1) * BR);
            compute(va, wa, row_c + b * BR, b * KS, hh, hl);
            load(va, wa, row_c + (b + 2) * BR);              // b + 2 == NB: first batch of the next chunk
            compute(vb, wb, row_c + (b + 1) * BR, (b + 1) * KS, hh, hl);
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 rv = run[(m * 4 + j) * 256 + threadIdx.x];
                rv.x += acc[m][j][0]; rv.y += acc[m][j][1]; rv.z += acc[m][j][2]; rv.w += acc[m][j][3];
                run[(m * 4 + j) * 256 + threadIdx.x] = rv;
            }
        if (c + 1 < nch) hstore((c + 1) & 1);
    }
    // ---- per-CTA partial: lane (g, t) owns rows c = g, g+8 (+16 m) and columns cb+8t .. cb+8t+7 ----
    float* pg = partG + (size_t)bid * R * d;
    const int oc = cb + 8 * t;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        float4 rj[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rj[j] = run[(m * 4 + j) * 256 + threadIdx.x];
        const float e[4][4] = {{rj[0].x, rj[0].y, rj[0].z, rj[0].w}, {rj[1].x, rj[1].y, rj[1].z, rj[1].w},
                               {rj[2].x, rj[2].y, rj[2].z, rj[2].w}, {rj[3].x, rj[3].y, rj[3].z, rj[3].w}};
#pragma unroll
        for (int half = 0; half < 2; ++half) {               // c = g (+8)
            const int c = m * 16 + g + 8 * half;
            const int i0 = 2 * half;                          // c0/c1 for row g, c2/c3 for row g+8
            if (oc < d)
                *reinterpret_cast<float4*>(pg + (size_t)c * d + oc) = make_float4(e[0][i0], e[1][i0], e[2][i0], e[3][i0]);
            if (oc + 4 < d)
                *reinterpret_cast<float4*>(pg + (size_t)c * d + oc + 4) =
                    make_float4(e[0][i0 + 1], e[1][i0 + 1], e[2][i0 + 1], e[3][i0 + 1]);
        }
    }
    csum = f4_add(csum, f4_shfl_xor(csum, 1));
    csum = f4_add(csum, f4_shfl_xor(csum, 2));
    if (partCol && t == 0 && col_ok) *reinterpret_cast<float4*>(partCol + (size_t)bid * d + col) = csum;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if (lane == 0) s_dot[warp] = dot;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (partDot) {
            float tsum = 0.f;
            for (int w_ = 0; w_ < 8; ++w_) tsum += s_dot[w_];
            if (col_base == 0) partDot[bid] = tsum; else partDot[bid] += tsum;   // launches are stream-ordered
        }
        if (bid == 0) header[header_slot] = nblocks;
    }
}

template <int R, bool HAS_B>
__global__ void __launch_bounds__(256, R == 16 ? 2 : 1)
k_wgrad_stream(const float* __restrict__ A, int64_t lda, const float* __restrict__ H, const float* __restrict__ B, int64_t ldb,
               float* __restrict__ partG, float* __restrict__ partCol, float* __restrict__ partDot, int* header, int header_slot,
               int n, int d, int col_base) {
    extern __shared__ __align__(16) uint4 smem_q[];
    wgrad_stream_body<R, HAS_B>(A, lda, H, B, ldb, partG, partCol, partDot, header, header_slot, n, d, col_base, blockIdx.x,
                                gridDim.x, smem_q);
}

template <int R>
int launch_wgrad_t(const float* A, int64_t lda, const float* H, const float* B, int64_t ldb, float* partG, float* partCol,
                   float* partDot, int* header, int slot, int n, int d, cudaStream_t st) {
    if constexpr (R == 16 || R == 32) {
        if (tc_enabled()) {
            const int ntiles = (n + kMmaRows - 1) / kMmaRows;
            int grid = ntiles < 1 ? 1 : ntiles;
            const int per_sm = R == 16 ? 2 : 1;
            const int cap = kMaxParts < per_sm * num_sms() ? kMaxParts : per_sm * num_sms();
            if (grid > cap) grid = cap;
            constexpr size_t smem = wgrad_stream_smem<R>();
            if (B) GCA_TRY(set_smem(k_wgrad_stream<R, true>, smem, true));
            else GCA_TRY(set_smem(k_wgrad_stream<R, false>, smem, true));
            for (int col0 = 0; col0 < d; col0 += 256) {
                ProfScope ps(slot == 0 ? "wgrad_up" : "wgrad_down", st, "mma");
                if (B)
                    GCA_CUDA(launch_pdl(k_wgrad_stream<R, true>, dim3(grid), dim3(256), smem, st, A, lda, H, B, ldb, partG, partCol, partDot, header, slot, n, d, col0));
                else
                    GCA_CUDA(launch_pdl(k_wgrad_stream<R, false>, dim3(grid), dim3(256), smem, st, A, lda, H, B, ldb, partG, partCol, partDot, header, slot, n, d, col0));
                count_launch();
            }
            GCA_CUDA(cudaGetLastError());
            return GCA_OK;
        }
    }
    constexpr int CW = R < 16 ? R : 16, CB = R / CW;
    constexpr int kMaxWarps = 12;
    const int max_chunks = kMaxWarps / CB;                 // column chunks (128 wide) per launch
    const int nblocks = (n + kWgRows - 1) / kWgRows;
    int grid = nblocks < 1 ? 1 : nblocks;
    const int cap = kMaxParts < 2 * num_sms() ? kMaxParts : 2 * num_sms();
    if (grid > cap) grid = cap;
    for (int col0 = 0; col0 < d; col0 += max_chunks * 128) {
        const int dsub = (d - col0) < max_chunks * 128 ? (d - col0) : max_chunks * 128;
        const int nchunks = (dsub / 4 + 31) / 32;
        int RS = kMaxWarps / (nchunks * CB);
        if (RS < 1) RS = 1;
        const int warps = nchunks * CB * RS;
        const size_t smem_h = sizeof(float) * (size_t)kWgRows * R;
        const size_t smem_g = sizeof(float) * ((size_t)R * dsub + dsub);
        const size_t smem = smem_h > smem_g ? smem_h : smem_g;
        if (smem > 200 * 1024) return GCA_ERR_UNSUPPORTED;
        GCA_TRY(set_smem(k_wgrad<R>, smem));
        {
            ProfScope ps(slot == 0 ? "wgrad_up" : "wgrad_down", st, "ffma");
            k_wgrad<R><<<grid, warps * 32, smem, st>>>(A, lda, H, B, ldb, partG, partCol, partDot, header, slot, n, d,
                                                        col0, dsub, nchunks, RS);
        }
        GCA_LAUNCH_OK();
    }
    return GCA_OK;
}

}  // namespace

int launch_wgrad(int r, const float* A, int64_t lda, const float* H, const float* B, int64_t ldb, float* partG, float* partCol,
                 float* partDot, int* header, int slot, int n, int d, cudaStream_t st) {
    GCA_DISPATCH_R(r, (launch_wgrad_t<R_>(A, lda, H, B, ldb, partG, partCol, partDot, header, slot, n, d, st)));
}

}  // namespace gca

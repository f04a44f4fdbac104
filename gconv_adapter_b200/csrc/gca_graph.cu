// K0 -- graph-structure build: COO int64 edge_index -> (CSR by target, CSR by source, dis).
//
// Replaces what torch_geometric's gcn_norm / add_remaining_self_loops recompute inside both
// GCNConv calls of every reference forward (/root/reference/src/finetune/gconv_adapter.py:92).
// Steps, all on the caller's stream, no host synchronisation:
//   count  : one pass over the E edges, integer atomics into per-row counters
//   scan   : exclusive scan -> rowptr (+1 per row for the added self loop), dis = 1/sqrt(deg)
//   fill   : second pass, atomic cursors place every neighbour id in its row
//   sort   : every row sorted ascending (normalised bitonic network) -> the layout, and
//            therefore every fp32 sum taken over it, is independent of atomic ordering
// Small graphs (the per-step rebuild of batched molecules) run all four steps in ONE
// single-CTA launch.
#include <new>

#include "gca_common.cuh"

namespace gca {

thread_local cudaError_t tl_last_cuda_error = cudaSuccess;

namespace {

constexpr int kSortSmemCap = 4096;   // ints of shared memory a CTA uses to sort one long row
constexpr int kSmallEdgeCap = 16384; // E + n at or below this: single-CTA fused build
constexpr int kSortRowsPerCta = 64;  // rows a CTA sorts per chunk (bounds the long-row list)

// ---------------- device pieces shared by the multi-launch and the fused build -------------
__device__ __forceinline__ void count_edges(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                            int64_t E, int N, int rb, int re, int normalize,
                                            int* cnt, int* cnt_t, int* flags, int64_t tid, int64_t nthreads) {
    for (int64_t e = tid; e < E; e += nthreads) {
        const int64_t s = src[e], t = dst[e];
        if ((uint64_t)s >= (uint64_t)N || (uint64_t)t >= (uint64_t)N) { flags[0] = 1; continue; }
        if (normalize && s == t) continue;               // add_remaining_self_loops drops these
        if (t >= rb && t < re) atomicAdd(&cnt[t - rb], 1);
        if (s >= rb && s < re) atomicAdd(&cnt_t[s - rb], 1);
    }
}

__device__ __forceinline__ void fill_edges(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                           int64_t E, int N, int rb, int re, int normalize,
                                           const int* __restrict__ rowptr, const int* __restrict__ rowptr_t,
                                           int* cur, int* cur_t, int* colidx, int* colidx_t,
                                           int64_t tid, int64_t nthreads) {
    const int n = re - rb;
    const int64_t total = E + (normalize ? n : 0);
    for (int64_t idx = tid; idx < total; idx += nthreads) {
        if (idx < E) {
            const int64_t s = src[idx], t = dst[idx];
            if ((uint64_t)s >= (uint64_t)N || (uint64_t)t >= (uint64_t)N) continue;
            if (normalize && s == t) continue;
            if (t >= rb && t < re) colidx[rowptr[t - rb] + atomicAdd(&cur[t - rb], 1)] = (int)s;
            if (s >= rb && s < re) colidx_t[rowptr_t[s - rb] + atomicAdd(&cur_t[s - rb], 1)] = (int)t;
        } else {                                          // the one (i, i) loop per node
            const int i = (int)(idx - E);
            colidx[rowptr[i] + atomicAdd(&cur[i], 1)] = i + rb;
            colidx_t[rowptr_t[i] + atomicAdd(&cur_t[i], 1)] = i + rb;
        }
    }
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive scan of (cnt[i] + add) by ONE CTA of 1024 threads; also dis (forward CSR only).
__device__ void scan_counts(const int* __restrict__ cnt, int n, int add, int* __restrict__ rowptr,
                            float* dis, int normalize, int* total_out, int* s_warp /*[33]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += blockDim.x * 4) {
        const int i0 = base + tid * 4;
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? cnt[i0 + k] + add : 0;
        const int tsum = v[0] + v[1] + v[2] + v[3];
        const int incl = warp_incl_scan(tsum, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = (lane < nwarps) ? s_warp[lane] : 0;
            const int wi = warp_incl_scan(w, lane);
            s_warp[lane] = wi - w;
            if (lane == 31) s_warp[32] = wi;
        }
        __syncthreads();
        int run = carry + s_warp[warp] + incl - tsum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < n) {
                rowptr[i0 + k] = run;
                // torch: deg.pow(-0.5) == 1.0f / sqrtf(deg) bit for bit
                if (dis) dis[i0 + k] = normalize ? __fdiv_rn(1.0f, __fsqrt_rn((float)v[k])) : 1.0f;
            }
            run += v[k];
        }
        carry += s_warp[32];
        __syncthreads();
    }
    if (tid == 0) { rowptr[n] = carry; *total_out = carry; }
}

// Normalised bitonic network on a[0..L) (shared or global), all comparators ascending, so the
// virtual +inf padding up to the next power of two never moves and is simply skipped.
__device__ void cta_sort(int* a, int L) {
    int P = 1;
    while (P < L) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        const int hk = k >> 1;
        for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
            const int i = (t / hk) * k + (t % hk);
            const int p = i ^ (k - 1);
            if (p < L) { const int x = a[i], y = a[p]; if (x > y) { a[i] = y; a[p] = x; } }
        }
        __syncthreads();
        for (int j = hk >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = 2 * j * (t / j) + (t % j);
                const int p = i + j;
                if (p < L) { const int x = a[i], y = a[p]; if (x > y) { a[i] = y; a[p] = x; } }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int warp_sort32(int v, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        {
            const int pv = __shfl_xor_sync(0xffffffffu, v, k - 1);
            v = ((lane & (k >> 1)) == 0) ? min(v, pv) : max(v, pv);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            const int pv = __shfl_xor_sync(0xffffffffu, v, j);
            v = ((lane & j) == 0) ? min(v, pv) : max(v, pv);
        }
    }
    return v;
}

// Sort the rows [row_lo, row_hi) of one CSR with the whole CTA: warps take the short rows,
// then the CTA walks the long ones together.
__device__ void sort_rows_cta(const int* __restrict__ rowptr, int* colidx, int row_lo, int row_hi, int* s_buf) {
    __shared__ int s_nlong;
    __shared__ int s_long[kSortRowsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int base = row_lo; base < row_hi; base += kSortRowsPerCta) {
        const int hi = min(row_hi, base + kSortRowsPerCta);
        if (threadIdx.x == 0) s_nlong = 0;
        __syncthreads();
        for (int row = base + warp; row < hi; row += nwarps) {
            const int beg = rowptr[row], L = rowptr[row + 1] - beg;
            if (L > 32) {
                if (lane == 0) s_long[atomicAdd(&s_nlong, 1)] = row;   // at most one entry per row of the chunk
            } else if (L > 1) {
                int v = (lane < L) ? colidx[beg + lane] : 0x7fffffff;
                v = warp_sort32(v, lane);
                if (lane < L) colidx[beg + lane] = v;
            }
        }
        __syncthreads();
        const int nlong = s_nlong;
        for (int q = 0; q < nlong; ++q) {                // uniform across the CTA; order is irrelevant
            const int row = s_long[q];
            const int beg = rowptr[row], L = rowptr[row + 1] - beg;
            if (L <= kSortSmemCap) {
                for (int i = threadIdx.x; i < L; i += blockDim.x) s_buf[i] = colidx[beg + i];
                __syncthreads();
                cta_sort(s_buf, L);
                for (int i = threadIdx.x; i < L; i += blockDim.x) colidx[beg + i] = s_buf[i];
                __syncthreads();
            } else {
                __syncthreads();
                cta_sort(colidx + beg, L);               // hub row: sort in place through L2
            }
        }
        __syncthreads();
    }
}

// List the hub rows of one CSR in row order (one CTA): hubitem[row] = first work item or -1, item_row[item] = row.
// Row order makes the item numbering - and with it every partial-sum order - independent of scheduling.
__device__ void list_hubs(const int* __restrict__ rowptr, int n, int* __restrict__ hubitem, int* __restrict__ item_row,
                          int* total_out, int* s_warp /*[33]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int row = base + tid;
        int items = 0;
        if (row < n) {
            const int deg = rowptr[row + 1] - rowptr[row];
            if (deg > kHubDeg) items = (deg + kHubChunk - 1) / kHubChunk;
        }
        const int incl = warp_incl_scan(items, lane);
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = (lane < nwarps) ? s_warp[lane] : 0;
            const int wi = warp_incl_scan(w, lane);
            s_warp[lane] = wi - w;
            if (lane == 31) s_warp[32] = wi;
        }
        __syncthreads();
        if (row < n) {
            const int first = carry + s_warp[warp] + incl - items;
            hubitem[row] = items ? first : -1;
            for (int c = 0; c < items; ++c) item_row[first + c] = row;
        }
        carry += s_warp[32];
        __syncthreads();
    }
    if (tid == 0) *total_out = carry;
}

__global__ void __launch_bounds__(1024) k0_hubs(const int* rowptr, const int* rowptr_t, int n, int* hubitem, int* hubitem_t,
                                                int* item_row, int* item_row_t, int* flags) {
    __shared__ int s_warp[33];
    if (blockIdx.x == 0) list_hubs(rowptr, n, hubitem, item_row, &flags[3], s_warp);
    else list_hubs(rowptr_t, n, hubitem_t, item_row_t, &flags[4], s_warp);
}

// ---------------- multi-launch build ---------------------------------------------------------
__global__ void k0_count(const int64_t* src, const int64_t* dst, int64_t E, int N, int rb, int re,
                         int normalize, int* cnt, int* cnt_t, int* flags) {
    count_edges(src, dst, E, N, rb, re, normalize, cnt, cnt_t, flags,
                (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(1024) k0_scan(const int* cnt, const int* cnt_t, int n, int normalize,
                                                int* rowptr, int* rowptr_t, float* dis, int* flags) {
    __shared__ int s_warp[33];
    if (blockIdx.x == 0) scan_counts(cnt, n, normalize ? 1 : 0, rowptr, dis, normalize, &flags[1], s_warp);
    else scan_counts(cnt_t, n, normalize ? 1 : 0, rowptr_t, nullptr, normalize, &flags[2], s_warp);
}

__global__ void k0_fill(const int64_t* src, const int64_t* dst, int64_t E, int N, int rb, int re, int normalize,
                        const int* rowptr, const int* rowptr_t, int* cur, int* cur_t, int* colidx, int* colidx_t) {
    fill_edges(src, dst, E, N, rb, re, normalize, rowptr, rowptr_t, cur, cur_t, colidx, colidx_t,
               (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

__global__ void __launch_bounds__(256) k0_sort(const int* rowptr, int* colidx, const int* rowptr_t, int* colidx_t, int n) {
    __shared__ int s_buf[kSortSmemCap];
    const int* rp = blockIdx.y ? rowptr_t : rowptr;
    int* ci = blockIdx.y ? colidx_t : colidx;
    const int nblk = (n + kSortRowsPerCta - 1) / kSortRowsPerCta;
    for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
        const int lo = b * kSortRowsPerCta;
        sort_rows_cta(rp, ci, lo, min(n, lo + kSortRowsPerCta), s_buf);
        __syncthreads();
    }
}

// ---------------- fused single-CTA build for small graphs ----------------------------------------
__global__ void __launch_bounds__(1024) k0_build_small(const int64_t* src, const int64_t* dst, int64_t E, int N,
                                                       int rb, int re, int normalize, int* cnt, int* cnt_t,
                                                       int* rowptr, int* rowptr_t, int* colidx, int* colidx_t,
                                                       float* dis, int* flags, int* hubitem, int* hubitem_t,
                                                       int* item_row, int* item_row_t) {
    __shared__ int s_warp[33];
    __shared__ int s_buf[kSortSmemCap];
    const int n = re - rb;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) { cnt[i] = 0; cnt_t[i] = 0; }
    if (tid < 8) flags[tid] = 0;
    __syncthreads();
    count_edges(src, dst, E, N, rb, re, normalize, cnt, cnt_t, flags, tid, nt);
    __syncthreads();
    scan_counts(cnt, n, normalize ? 1 : 0, rowptr, dis, normalize, &flags[1], s_warp);
    __syncthreads();
    scan_counts(cnt_t, n, normalize ? 1 : 0, rowptr_t, nullptr, normalize, &flags[2], s_warp);
    __syncthreads();
    for (int i = tid; i < n; i += nt) { cnt[i] = 0; cnt_t[i] = 0; }
    __syncthreads();
    fill_edges(src, dst, E, N, rb, re, normalize, rowptr, rowptr_t, cnt, cnt_t, colidx, colidx_t, tid, nt);
    __syncthreads();
    sort_rows_cta(rowptr, colidx, 0, n, s_buf);
    __syncthreads();
    sort_rows_cta(rowptr_t, colidx_t, 0, n, s_buf);
    __syncthreads();
    list_hubs(rowptr, n, hubitem, item_row, &flags[3], s_warp);
    __syncthreads();
    list_hubs(rowptr_t, n, hubitem_t, item_row_t, &flags[4], s_warp);
}

__global__ void k0_edge_coef(const int* __restrict__ rowptr, const int* __restrict__ colidx,
                             const float* __restrict__ dis, float* coef, int n) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int row = warp; row < n; row += nwarps) {
        const int beg = rowptr[row], end = rowptr[row + 1];
        const float di = dis[row];
        for (int e = beg + lane; e < end; e += 32) coef[e] = __fmul_rn(dis[colidx[e]], di);   // dis[src] * 1 * dis[dst]
    }
}

struct Layout {
    size_t rowptr, rowptr_t, colidx, colidx_t, dis, hubitem, hubitem_t, item_row, item_row_t, cnt, cnt_t, flags, total;
    size_t hub_cap;
};
Layout make_layout(int64_t E, int32_t n) {
    Layout L;
    size_t off = 0;
    const size_t cap = (size_t)E + (size_t)n;
    L.rowptr = off;   off += align_up(sizeof(int32_t) * ((size_t)n + 1));
    L.rowptr_t = off; off += align_up(sizeof(int32_t) * ((size_t)n + 1));
    L.colidx = off;   off += align_up(sizeof(int32_t) * (cap + 1));
    L.colidx_t = off; off += align_up(sizeof(int32_t) * (cap + 1));
    L.dis = off;      off += align_up(sizeof(float) * ((size_t)n + 1));
    // hub items: sum over hub rows of ceil(deg / kHubChunk) <= 2 * entries / kHubChunk + 1
    L.hub_cap = 2 * cap / kHubChunk + 2;
    L.hubitem = off;    off += align_up(sizeof(int32_t) * ((size_t)n + 1));
    L.hubitem_t = off;  off += align_up(sizeof(int32_t) * ((size_t)n + 1));
    L.item_row = off;   off += align_up(sizeof(int32_t) * L.hub_cap);
    L.item_row_t = off; off += align_up(sizeof(int32_t) * L.hub_cap);
    L.cnt = off;      off += align_up(sizeof(int32_t) * ((size_t)n + 1));
    L.cnt_t = off;    off += align_up(sizeof(int32_t) * ((size_t)n + 1));
    L.flags = off;    off += align_up(sizeof(int32_t) * 8);
    L.total = off;
    return L;
}

}  // namespace
}  // namespace gca

using namespace gca;

extern "C" size_t gca_graph_workspace_bytes(int64_t E, int32_t N, int32_t row_begin, int32_t row_end) {
    if (E < 0 || N < 0 || row_begin < 0 || row_end < row_begin || row_end > N) return 0;
    return make_layout(E, row_end - row_begin).total;
}

extern "C" int gca_graph_build(const int64_t* src, const int64_t* dst, int64_t E, int32_t N,
                               int32_t row_begin, int32_t row_end, int normalize,
                               void* workspace, size_t workspace_bytes, gca_stream_t stream_, gca_graph** out) {
    if (!out) return GCA_ERR_INVALID_ARG;
    *out = nullptr;
    if (E < 0 || N < 0 || row_begin < 0 || row_end < row_begin || row_end > N) return GCA_ERR_INVALID_ARG;
    if (E > 0 && (!src || !dst)) return GCA_ERR_INVALID_ARG;
    const int32_t n = row_end - row_begin;
    if ((int64_t)E + n >= (int64_t)0x7fffffff) return GCA_ERR_UNSUPPORTED;    // int32 CSR
    const Layout L = make_layout(E, n);
    if (!workspace || workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(workspace) % kAlign) != 0)
        return GCA_ERR_WORKSPACE;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    char* base = static_cast<char*>(workspace);

    gca_graph* g = new (std::nothrow) gca_graph();
    if (!g) return GCA_ERR_INVALID_ARG;
    g->N = N; g->row_begin = row_begin; g->row_end = row_end; g->normalize = normalize ? 1 : 0;
    g->E = E; g->capacity = E + n;
    g->rowptr = reinterpret_cast<int32_t*>(base + L.rowptr);
    g->rowptr_t = reinterpret_cast<int32_t*>(base + L.rowptr_t);
    g->colidx = reinterpret_cast<int32_t*>(base + L.colidx);
    g->colidx_t = reinterpret_cast<int32_t*>(base + L.colidx_t);
    g->dis = reinterpret_cast<float*>(base + L.dis);
    g->hubitem = reinterpret_cast<int32_t*>(base + L.hubitem);
    g->hubitem_t = reinterpret_cast<int32_t*>(base + L.hubitem_t);
    g->item_row = reinterpret_cast<int32_t*>(base + L.item_row);
    g->item_row_t = reinterpret_cast<int32_t*>(base + L.item_row_t);
    g->hub_cap = (int64_t)L.hub_cap;
    g->nitems = g->nitems_t = -1;
    g->cnt = reinterpret_cast<int32_t*>(base + L.cnt);
    g->cnt_t = reinterpret_cast<int32_t*>(base + L.cnt_t);
    g->flags = reinterpret_cast<int32_t*>(base + L.flags);

    auto fail = [&](int st) { delete g; return st; };
    const int norm = g->normalize;

    if (E + n <= kSmallEdgeCap) {
        k0_build_small<<<1, 1024, 0, stream>>>(src, dst, E, N, row_begin, row_end, norm, g->cnt, g->cnt_t,
                                               g->rowptr, g->rowptr_t, g->colidx, g->colidx_t, g->dis, g->flags,
                                               g->hubitem, g->hubitem_t, g->item_row, g->item_row_t);
        count_launch();
        if (record(cudaGetLastError()) != GCA_OK) return fail(GCA_ERR_CUDA);
        *out = g;
        return GCA_OK;
    }

    // cnt, cnt_t and flags are contiguous in the layout: one memset clears them.
    if (record(cudaMemsetAsync(g->cnt, 0, L.total - L.cnt, stream)) != GCA_OK) return fail(GCA_ERR_CUDA);
    const int threads = 256;
    const int sms = num_sms();
    const int64_t want = (E + threads - 1) / threads;
    const int grid_e = (int)(want < 1 ? 1 : (want > (int64_t)sms * 16 ? (int64_t)sms * 16 : want));
    k0_count<<<grid_e, threads, 0, stream>>>(src, dst, E, N, row_begin, row_end, norm, g->cnt, g->cnt_t, g->flags);
    count_launch();
    k0_scan<<<2, 1024, 0, stream>>>(g->cnt, g->cnt_t, n, norm, g->rowptr, g->rowptr_t, g->dis, g->flags);
    count_launch();
    if (record(cudaMemsetAsync(g->cnt, 0, L.flags - L.cnt, stream)) != GCA_OK) return fail(GCA_ERR_CUDA);
    k0_fill<<<grid_e, threads, 0, stream>>>(src, dst, E, N, row_begin, row_end, norm, g->rowptr, g->rowptr_t,
                                            g->cnt, g->cnt_t, g->colidx, g->colidx_t);
    count_launch();
    const int nblk = (n + kSortRowsPerCta - 1) / kSortRowsPerCta;
    dim3 grid_s((unsigned)(nblk < 1 ? 1 : (nblk > sms * 8 ? sms * 8 : nblk)), 2);
    k0_sort<<<grid_s, 256, 0, stream>>>(g->rowptr, g->colidx, g->rowptr_t, g->colidx_t, n);
    count_launch();
    k0_hubs<<<2, 1024, 0, stream>>>(g->rowptr, g->rowptr_t, n, g->hubitem, g->hubitem_t, g->item_row, g->item_row_t, g->flags);
    count_launch();
    if (record(cudaGetLastError()) != GCA_OK) return fail(GCA_ERR_CUDA);
    *out = g;
    return GCA_OK;
}

extern "C" int gca_graph_validate(gca_graph* g, gca_stream_t stream_, int64_t* nnz, int64_t* nnz_t) {
    if (!g) return GCA_ERR_INVALID_ARG;
    int32_t h[5] = {0, 0, 0, 0, 0};
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    GCA_CUDA(cudaMemcpyAsync(h, g->flags, sizeof(h), cudaMemcpyDeviceToHost, stream));
    GCA_CUDA(cudaStreamSynchronize(stream));
    if (nnz) *nnz = h[1];
    if (nnz_t) *nnz_t = h[2];
    g->nitems = h[3];
    g->nitems_t = h[4];
    return h[0] ? GCA_ERR_INDEX_RANGE : GCA_OK;
}

// Deferred validation: gca_graph_validate synchronises the stream on every build - once per step when the graph changes
// every step (batched molecules).  The asynchronous pair copies the flags into caller-owned PINNED host memory (5 x int32)
// on the stream; the caller polls an event it recorded after this call and then lets _finish read them (host only).
extern "C" int gca_graph_validate_async(const gca_graph* g, int32_t* host_flags5, gca_stream_t stream_) {
    if (!g || !host_flags5) return GCA_ERR_INVALID_ARG;
    GCA_CUDA(cudaMemcpyAsync(host_flags5, g->flags, 5 * sizeof(int32_t), cudaMemcpyDeviceToHost, static_cast<cudaStream_t>(stream_)));
    return GCA_OK;
}
extern "C" int gca_graph_validate_finish(gca_graph* g, const int32_t* host_flags5, int64_t* nnz, int64_t* nnz_t) {
    if (!g || !host_flags5) return GCA_ERR_INVALID_ARG;
    if (nnz) *nnz = host_flags5[1];
    if (nnz_t) *nnz_t = host_flags5[2];
    g->nitems = host_flags5[3];
    g->nitems_t = host_flags5[4];
    return host_flags5[0] ? GCA_ERR_INDEX_RANGE : GCA_OK;
}

extern "C" void gca_graph_destroy(gca_graph* g) { delete g; }

extern "C" int gca_graph_get_view(const gca_graph* g, gca_graph_view* v) {
    if (!g || !v) return GCA_ERR_INVALID_ARG;
    v->N = g->N; v->row_begin = g->row_begin; v->row_end = g->row_end; v->normalize = g->normalize;
    v->capacity = g->capacity;
    v->rowptr = g->rowptr; v->colidx = g->colidx; v->rowptr_t = g->rowptr_t; v->colidx_t = g->colidx_t;
    v->dis = g->dis;
    return GCA_OK;
}

extern "C" int gca_graph_edge_coef(const gca_graph* g, float* coef, gca_stream_t stream_) {
    if (!g || !coef) return GCA_ERR_INVALID_ARG;
    if (g->row_begin != 0 || g->row_end != g->N) return GCA_ERR_UNSUPPORTED;
    const int n = g->N;
    if (n == 0) return GCA_OK;
    const int grid = (n + 7) / 8 < num_sms() * 8 ? (n + 7) / 8 : num_sms() * 8;
    k0_edge_coef<<<grid < 1 ? 1 : grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(g->rowptr, g->colidx, g->dis, coef, n);
    GCA_LAUNCH_OK();
    return GCA_OK;
}

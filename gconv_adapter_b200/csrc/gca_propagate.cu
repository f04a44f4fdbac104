#include "gca_common.cuh"
#include "gca_device.cuh"
#include "gca_host.cuh"

using namespace gca;

// ------------------------------------------------------------------------------------------
// d-wide propagation over the SAME graph handle (the backbone's normalised SpMM):
//     out[i, :] = dis[i] * sum_{j in N(i)} dis[j] * X[j, :]
// i.e. DIFFormer's gcn_conv (src/models/transductive/difformer.py:63-79) and the aggregation inside NodeFormer's
// add_conv_relational_bias (nodeformer.py:202-224) for a graph that already holds one self loop per node (the
// transductive pipeline adds them: scripts/finetune_transductive_learning.py:112-113), with all heads flattened
// into the column dimension.  One warp per (row, 128-column block): the lanes own one float4 of the block, the
// neighbour ids and their dis are read once per warp (lanes 0-7 of a batch of 8, broadcast by shuffle), 8 neighbour
// rows are in flight per lane; sums run in ascending neighbour order (deterministic).  transpose = 1 walks the
// CSR by source: the adjoint, used for the backward.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_propagate(const int* __restrict__ rowptr, const int* __restrict__ colidx, const float* __restrict__ nbr_scale,
            const float* __restrict__ row_scale, const float* __restrict__ X, int64_t ldx, float* __restrict__ out, int64_t ldo,
            int n, int D) {
    const int lane = threadIdx.x & 31;
    const int nblk = (D + 127) / 128;
    const long long wglobal = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long wtotal = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long item = wglobal; item < (long long)n * nblk; item += wtotal) {
        const int row = (int)(item / nblk), cb = (int)(item - (long long)row * nblk);
        const int col = cb * 128 + lane * 4;
        const bool col_ok = col < D;                                  // D % 4 == 0
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int e = beg; e < end; e += 8) {
            int jm = -1;
            float dm = 0.f;
            if (lane < 8 && e + lane < end) { jm = __ldg(colidx + e + lane); dm = __ldg(nbr_scale + jm); }
            float4 v[8];
            float w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = __shfl_sync(0xffffffffu, jm, u);
                w[u] = __shfl_sync(0xffffffffu, dm, u);
                v[u] = (j >= 0 && col_ok) ? ldg4(X + (size_t)j * ldx + col) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y);
                acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w);
            }
        }
        if (col_ok) {
            const float di = __ldg(row_scale + row);
            *reinterpret_cast<float4*>(out + (size_t)row * ldo + col) = f4_scale(acc, di);
        }
    }
}

extern "C" int gca_propagate(const gca_graph* g, int transpose, const float* X_full, int64_t ldx, float* out_local, int64_t ldo,
                             const float* src_scale, const float* dst_scale, int32_t D, gca_stream_t stream) {
    if (!g || !X_full || !out_local || D <= 0 || (D % 4) != 0 || ldx < D || ldo < D || (ldx % 4) != 0 || (ldo % 4) != 0)
        return GCA_ERR_INVALID_ARG;
    if (g->row_begin != 0 || g->row_end != g->N) return GCA_ERR_UNSUPPORTED;   // needs dis of every neighbour: full-graph handles only
    const int n = g->N;
    if (n == 0) return GCA_OK;
    const long long items = (long long)n * ((D + 127) / 128);
    long long grid = (items + 7) / 8;
    if (grid > 16LL * num_sms()) grid = 16LL * num_sms();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        ProfScope ps(transpose ? "propagate_t" : "propagate", st);
        // forward: rows = targets i (scaled by dst_scale), neighbours = sources j (src_scale); the adjoint walks the CSR by
        // source: rows = sources, neighbours = targets, the two scales swap roles
        const float* ss = src_scale ? src_scale : g->dis;
        const float* ds = dst_scale ? dst_scale : g->dis;
        k_propagate<<<(int)grid, 256, 0, st>>>(transpose ? g->rowptr_t : g->rowptr, transpose ? g->colidx_t : g->colidx,
                                               transpose ? ds : ss, transpose ? ss : ds, X_full, ldx, out_local, ldo, n, D);
    }
    GCA_LAUNCH_OK();
    return GCA_OK;
}

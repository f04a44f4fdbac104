// Issue-rate micro-benchmark of the legacy tensor path on B200: mma.sync m16n8k8 tf32 vs m16n8k16 f16 (fp32 accumulate).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate experiments/mma_rate.cu && ./mma_rate
#include <cuda_runtime.h>
#include <cstdio>
#include <stdint.h>

template <int KIND, int CHAINS>
__global__ void k(float* out, int iters) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3c003c00u, b1 = 0x3c003c00u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
        }
    }
    float s = 0;
    for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int CHAINS>
void run(const char* name, int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, (size_t)sms * 1024 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, CHAINS><<<sms, warps_per_sm * 32>>>(out, 100);
    cudaEventRecord(e0);
    k<KIND, CHAINS><<<sms, warps_per_sm * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)sms * warps_per_sm * iters * CHAINS;
    const double flop = mmas * (KIND == 0 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16);
    printf("%-28s warps/SM %2d chains %d: %.1f TFLOP/s, %.2f clk per MMA per sub-core (at 1.965 GHz)\n", name, warps_per_sm, CHAINS,
           flop / ms / 1e9, ms * 1e-3 * 1.965e9 / (mmas / sms / 4));
    cudaFree(out);
}

int main() {
    run<0, 8>("tf32 m16n8k8", 4); run<0, 8>("tf32 m16n8k8", 8); run<0, 8>("tf32 m16n8k8", 16);
    run<1, 8>("f16  m16n8k16", 4); run<1, 8>("f16  m16n8k16", 8); run<1, 8>("f16  m16n8k16", 16);
    run<0, 2>("tf32 m16n8k8", 4); run<1, 2>("f16  m16n8k16", 4);
    return 0;
}

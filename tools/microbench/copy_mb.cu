// Row-copy ceilings: HBM -> 48 MB ring, one warp per 8 KB row, 64-bit vs 128-bit accesses, plain vs .cg stores
#include <cstdio>
#include <cuda_runtime.h>
constexpr int L = 1024;
template <int W, int CG>  // W = 8 or 16 bytes per lane per access
__global__ void __launch_bounds__(256, 2) k_copy(const char* __restrict__ in, char* __restrict__ ring, int nrows, int ring_rows) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + w; row < nrows; row += gridDim.x * 8) {
        const char* src = in + (size_t)row * L * 8;
        char* dst = ring + (size_t)(row % ring_rows) * L * 8;
        if (W == 8) {
            float2 v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __ldcg((const float2*)src + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < 32; ++i) { if (CG) __stcg((float2*)dst + lane + 32 * i, v[i]); else ((float2*)dst)[lane + 32 * i] = v[i]; }
        } else {
            float4 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __ldcg((const float4*)src + lane + 32 * i);
#pragma unroll
            for (int i = 0; i < 16; ++i) { if (CG) __stcg((float4*)dst + lane + 32 * i, v[i]); else ((float4*)dst)[lane + 32 * i] = v[i]; }
        }
    }
}
int main() {
    const int nimg = 192, ring_imgs = 6;
    char *in, *ring;
    cudaMalloc(&in, (size_t)nimg * L * L * 8); cudaMalloc(&ring, (size_t)ring_imgs * L * L * 8);
    cudaMemset(in, 0, (size_t)nimg * L * L * 8); cudaMemset(ring, 0, (size_t)ring_imgs * L * L * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](auto f, const char* name) {
        f(); cudaDeviceSynchronize();
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-40s %.3f ms = %.2f us/image  (%.2f TB/s read + same write)\n", name, ms, ms * 1e3 / nimg, nimg * 8.388608e6 / ms / 1e9);
    };
    for (int g : {148, 296, 592}) {
        printf("grid %d x 256\n", g);
        timeit([&] { k_copy<8, 1><<<g, 256>>>(in, ring, nimg * L, ring_imgs * L); }, "64-bit  HBM->ring  st.cg");
        timeit([&] { k_copy<8, 0><<<g, 256>>>(in, ring, nimg * L, ring_imgs * L); }, "64-bit  HBM->ring  st");
        timeit([&] { k_copy<16, 1><<<g, 256>>>(in, ring, nimg * L, ring_imgs * L); }, "128-bit HBM->ring  st.cg");
        timeit([&] { k_copy<8, 1><<<g, 256>>>(ring, ring, ring_imgs * L, ring_imgs * L); }, "64-bit  ring->ring (6 img only)");
        timeit([&] { k_copy<8, 1><<<g, 256>>>(in, in + (size_t)96 * L * L * 8, 96 * L, 96 * L); }, "64-bit  HBM->HBM (96 img)");
    }
    return 0;
}

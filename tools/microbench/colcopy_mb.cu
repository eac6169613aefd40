// Column-slab copy ceilings on an L2-resident 48 MB ring: segment width 64 B (8 columns) vs 128 B (16 columns) vs 256 B
#include <cstdio>
#include <cuda_runtime.h>
constexpr int L = 1024;
template <int CC>   // threads = CC * 32 ; thread (c = t % CC, tl = t / CC) moves rows tl + 32 i of column c
__global__ void __launch_bounds__(CC * 32) k_colcopy(float2* __restrict__ ring, int nslabs_total, int ring_imgs) {
    const int t = threadIdx.x, c = t % CC, tl = t / CC;
    constexpr int NS = L / CC;
    for (int s = blockIdx.x; s < nslabs_total; s += gridDim.x) {
        const int img = (s / NS) % ring_imgs, sl = s % NS;
        float2* base = ring + (size_t)img * L * L + sl * CC + c;
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __ldcg(base + (size_t)(tl + 32 * i) * L);
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i].x += 1.f; __stcg(base + (size_t)(tl + 32 * i) * L, v[i]); }
    }
}
int main() {
    const int nimg = 192, ring_imgs = 6;
    float2* ring; cudaMalloc(&ring, (size_t)ring_imgs * L * L * 8); cudaMemset(ring, 0, (size_t)ring_imgs * L * L * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](auto f, const char* name) {
        f(); cudaDeviceSynchronize();
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-32s %.3f ms = %.2f us/image (%.1f TB/s r+w) %s\n", name, ms, ms * 1e3 / nimg, nimg * 16.777e6 / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    timeit([&] { k_colcopy<4><<<148 * 4, 128>>>(ring, nimg * 256, ring_imgs); }, "CC=4  (32 B)  4 CTA/SM");
    timeit([&] { k_colcopy<8><<<148 * 2, 256>>>(ring, nimg * 128, ring_imgs); }, "CC=8  (64 B)  2 CTA/SM");
    timeit([&] { k_colcopy<8><<<148 * 4, 256>>>(ring, nimg * 128, ring_imgs); }, "CC=8  (64 B)  4 CTA/SM");
    timeit([&] { k_colcopy<16><<<148 * 1, 512>>>(ring, nimg * 64, ring_imgs); }, "CC=16 (128 B) 1 CTA/SM");
    timeit([&] { k_colcopy<16><<<148 * 2, 512>>>(ring, nimg * 64, ring_imgs); }, "CC=16 (128 B) 2 CTA/SM");
    timeit([&] { k_colcopy<32><<<148 * 1, 1024>>>(ring, nimg * 32, ring_imgs); }, "CC=32 (256 B) 1 CTA/SM");
    return 0;
}

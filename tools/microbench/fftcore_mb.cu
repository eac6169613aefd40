// Compute-only ceiling of the 32-point-per-thread column pipeline (no global traffic): how fast can an SM run
// radix-32 -> exchange -> radix-32(table) -> inverse radix-32 -> exchange -> radix-32(table)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../style_transfer_based_holographic_imaging_b200/csrc -o fftcore_mb fftcore_mb.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_core.cuh"
using namespace asmb;
constexpr int CC = 8, L = 1024, ROWS = L + L / 32, TW = 31 * 32;

template <int REGCAP_BLOCKS>
__global__ void __launch_bounds__(256, REGCAP_BLOCKS) k_cols_only(float2* out, const float2* twg, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* slab = reinterpret_cast<float2*>(smem_raw);
    float2* tw = slab + ROWS * CC;
    const int t = threadIdx.x, c = t % CC, tl = t / CC;
    for (int i = t; i < TW; i += 256) tw[i] = twg[i];
    for (int i = t; i < ROWS * CC; i += 256) slab[i] = make_float2(i * 1e-4f, -i * 2e-4f);
    __syncthreads();
    using LAY = ColLayout32<CC>;
    float2* col = slab + c;
    float2 v[32];
    lds16<LAY, 5>(v, col + tl * CC);
    for (int r = 0; r < reps; ++r) {
        fwd32_first(v);
        sts16<LAY, 5>(v, col + tl * CC);
        __syncthreads();
        lds16<LAY, 0>(v, col + 33 * tl * CC);
        fwd32_table(v, tw + tl);
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i].x *= 0.03125f; v[i].y *= 0.03125f; }
        inv32_first(v);
        sts16<LAY, 0>(v, col + 33 * tl * CC);
        __syncthreads();
        lds16<LAY, 5>(v, col + tl * CC);
        inv32_table(v, tw + tl);
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i].x *= 0.03125f; v[i].y *= 0.03125f; }
    }
    float2 acc = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc.x += v[i].x; acc.y += v[i].y; }
    out[blockIdx.x * 256 + t] = acc;
}

// one warp per row variant (no CTA barrier): row pipeline fwd + inv
__global__ void __launch_bounds__(256, 2) k_rows_only(float2* out, const float2* twg, int reps) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* lines = reinterpret_cast<float2*>(smem_raw);
    float2* tw = lines + 8 * RowLayout32::line_elems(L);
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < TW; i += 256) tw[i] = twg[i];
    __syncthreads();
    float2* line = lines + w * RowLayout32::line_elems(L);
    float2 v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = make_float2(lane * 1e-3f + i, i * 1e-2f - lane);
    for (int r = 0; r < reps; ++r) {
        fwd32_first(v);
        sts16<RowLayout32, 5>(v, line + lane);
        __syncwarp();
        lds16<RowLayout32, 0>(v, line + 33 * lane);
        fwd32_table(v, tw + lane);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i].x *= 0.03125f; v[i].y *= 0.03125f; }
    }
    float2 acc = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 32; ++i) { acc.x += v[i].x; acc.y += v[i].y; }
    out[blockIdx.x * 256 + t] = acc;
}

int main() {
    float2 *out, *tw;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float2));
    cudaMalloc(&tw, TW * sizeof(float2));
    cudaMemset(tw, 0, TW * sizeof(float2));
    const size_t smem = (size_t)ROWS * CC * 8 + TW * 8;
    cudaFuncSetAttribute(k_cols_only<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_cols_only<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_rows_only, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int reps = 200;
    for (int ctas : {1, 2, 3}) {
        auto run = [&] { if (ctas == 1) k_cols_only<1><<<148 * ctas, 256, smem>>>(out, tw, reps); else k_cols_only<2><<<148 * ctas, 256, smem>>>(out, tw, reps); };
        run(); cudaDeviceSynchronize();
        cudaEventRecord(a); run(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        // one rep = a full column pass (fwd+inv FFT) of a slab of 8192 points per CTA
        double slabs = 148.0 * ctas * reps;
        printf("cols-only  %d CTA/SM: %.3f ms, %.2f us per 128 slabs (= one image column pass)  err=%s\n", ctas, ms, ms * 1e3 / slabs * 128, cudaGetErrorString(cudaGetLastError()));
    }
    for (int ctas : {1, 2}) {
        k_rows_only<<<148 * ctas, 256, smem>>>(out, tw, reps); cudaDeviceSynchronize();
        cudaEventRecord(a); k_rows_only<<<148 * ctas, 256, smem>>>(out, tw, reps); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double rows = 148.0 * ctas * 8 * reps;
        printf("rows-only  %d CTA/SM: %.3f ms, %.2f us per 1024 rows (= one image row pass, one direction)\n", ctas, ms, ms * 1e3 / rows * 1024);
    }
    return 0;
}

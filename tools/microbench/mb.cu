// Machine-characterisation microbenchmarks for the ASM propagator design (B200, sm_100a).
// Not part of the product; results are summarised in profiles/ and DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb mb.cu && ./mb
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

__global__ void k_ffma(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fadd(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = x[i] + ((i & 1) ? a : b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mix of FADD and FMUL with distinct register operands (like a butterfly)
__global__ void k_fadd3(float* out, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x + i + a;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            float p = x[i] + x[i + 1];
            float q = x[i] - x[i + 1];
            x[i] = p; x[i + 1] = q * b;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pk(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x[16];
    unsigned long long A = pk(a, a), B = pk(b, b);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { float lo, hi; upk(x[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_fadd2(float* out, float a, float b) {
    unsigned long long x[16];
    unsigned long long A = pk(a, b);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = pk(threadIdx.x + i, threadIdx.x - i);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(A));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { float lo, hi; upk(x[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed butterfly: p = x+y ; q = x-y on float2 pairs, distinct registers
__global__ void k_bfly2(float* out, float a, float b) {
    unsigned long long x[16];
    unsigned long long NEG = pk(-1.f, -1.f);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = pk(threadIdx.x + i + a, threadIdx.x - i + b);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            unsigned long long p, q;
            asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(x[i]), "l"(x[i + 1]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(q) : "l"(x[i + 1]), "l"(NEG), "l"(x[i]));
            x[i] = p; x[i + 1] = q;
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) { float lo, hi; upk(x[i], lo, hi); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(float* out, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

__global__ void k_mufu(float* out, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = __sinf(x[i]) + a;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared memory LDS.64 + STS.64 throughput, conflict-free
__global__ void k_smem(float* out) {
    extern __shared__ float2 sm[];
    int t = threadIdx.x;
    for (int i = t; i < 8192; i += blockDim.x) sm[i] = make_float2(i, -i);
    __syncthreads();
    float2 acc = make_float2(0, 0);
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float2 v = sm[(t + k * 512 + it) & 8191];
            acc.x += v.x; acc.y += v.y;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) sm[(t + k * 512) & 8191] = acc;
    }
    out[blockIdx.x * blockDim.x + t] = acc.x + acc.y;
}

// streaming read of a buffer (float4), repeated `reps` times: measures L2 (small buffer) or HBM (large)
__global__ void k_read(const float4* __restrict__ p, size_t n4, int reps, float* out) {
    float4 acc = make_float4(0, 0, 0, 0);
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++) {
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride * 4) {
            float4 v0 = p[i];
            float4 v1 = (i + stride < n4) ? p[i + stride] : make_float4(0, 0, 0, 0);
            float4 v2 = (i + 2 * stride < n4) ? p[i + 2 * stride] : make_float4(0, 0, 0, 0);
            float4 v3 = (i + 3 * stride < n4) ? p[i + 3 * stride] : make_float4(0, 0, 0, 0);
            acc.x += v0.x + v1.x + v2.x + v3.x; acc.y += v0.y + v1.y + v2.y + v3.y;
            acc.z += v0.z + v1.z + v2.z + v3.z; acc.w += v0.w + v1.w + v2.w + v3.w;
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

__global__ void k_write(float4* __restrict__ p, size_t n4, int reps, float v) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride)
            p[i] = make_float4(v, v + r, v, v);
}

// in-place read-modify-write (like the column pass on the L2-resident workspace)
__global__ void k_rmw(float4* __restrict__ p, size_t n4, int reps) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride * 4) {
            float4 v[4];
#pragma unroll
            for (int k = 0; k < 4; k++) if (i + k * stride < n4) v[k] = p[i + k * stride];
#pragma unroll
            for (int k = 0; k < 4; k++) if (i + k * stride < n4) { v[k].x += 1.f; p[i + k * stride] = v[k]; }
        }
}

// column-slab style access: each CTA reads 128-byte chunks with a row stride (like pass 2), L2 resident
__global__ void k_slab(const float4* __restrict__ p, int M /*row len in complex*/, int nimg, int reps, float* out) {
    // image = M rows x M complex (8B) ; slab = 16 complex = 128 B = 8 float4 ; M/16 slabs per image
    float4 acc = make_float4(0, 0, 0, 0);
    int slabs = M / 16;
    int row_f4 = M / 2;  // float4 per row
    for (int r = 0; r < reps; r++)
        for (int w = blockIdx.x; w < nimg * slabs; w += gridDim.x) {
            int img = w / slabs, s = w % slabs;
            const float4* base = p + (size_t)img * M * row_f4 + s * 8;
            for (int i = threadIdx.x; i < M * 8; i += blockDim.x * 4) {
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { int j = i + k * blockDim.x; v[k] = (j < M * 8) ? base[(size_t)(j >> 3) * row_f4 + (j & 7)] : make_float4(0, 0, 0, 0); }
#pragma unroll
                for (int k = 0; k < 4; k++) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
            }
        }
    if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

template <typename F>
float timeit(F f, int n = 3) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < n; i++) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s sms=%d smem/block optin=%zu smem/sm=%zu L2=%d MB regs/sm=%d clock=%d kHz\n", prop.name, sms,
           prop.sharedMemPerBlockOptin, prop.sharedMemPerMultiprocessor, prop.l2CacheSize >> 20, prop.regsPerMultiprocessor, prop.clockRate);
    int v; cudaDeviceGetAttribute(&v, cudaDevAttrMaxPersistingL2CacheSize, 0); printf("max persisting L2 = %d MB\n", v >> 20);
    float* out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024));
    int blocks = sms * 4, threads = 256;  // 1024 thr/SM = 32 warps
    double lanes = (double)blocks * threads;
    {
        float ms = timeit([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("FFMA        : %.1f G lane-instr/s  (%.2f /clk/SM @1.965GHz)\n", lanes * ITERS * 16 / ms / 1e6, lanes * ITERS * 16 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_fadd<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("FADD        : %.1f G lane-instr/s  (%.2f /clk/SM)\n", lanes * ITERS * 16 / ms / 1e6, lanes * ITERS * 16 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_fadd3<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("FADD/FMUL 3r: %.1f G lane-instr/s  (%.2f /clk/SM)\n", lanes * ITERS * 24 / ms / 1e6, lanes * ITERS * 24 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("FFMA2       : %.1f G lane-instr/s  (%.2f instr/clk/SM = %.2f fma/clk/SM)\n", lanes * ITERS * 16 / ms / 1e6, lanes * ITERS * 16 / ms / 1e6 / sms / 1.965, 2 * lanes * ITERS * 16 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_fadd2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("FADD2       : %.1f G lane-instr/s  (%.2f instr/clk/SM = %.2f add/clk/SM)\n", lanes * ITERS * 16 / ms / 1e6, lanes * ITERS * 16 / ms / 1e6 / sms / 1.965, 2 * lanes * ITERS * 16 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_bfly2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
        printf("BFLY2 (3reg): %.1f G lane-instr/s  (%.2f instr/clk/SM = %.2f flop-lanes/clk/SM)\n", lanes * ITERS * 16 / ms / 1e6, lanes * ITERS * 16 / ms / 1e6 / sms / 1.965, 2 * lanes * ITERS * 16 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_dfma<<<blocks, threads>>>(out, 1.0001, 0.5); });
        printf("DFMA        : %.1f G lane-instr/s  (%.2f /clk/SM)\n", lanes * ITERS * 8 / ms / 1e6, lanes * ITERS * 8 / ms / 1e6 / sms / 1.965);
        ms = timeit([&] { k_mufu<<<blocks, threads>>>(out, 0.5f); });
        printf("MUFU.SIN(+2): %.1f G lane-instr/s  (%.2f /clk/SM)\n", lanes * ITERS * 8 / ms / 1e6, lanes * ITERS * 8 / ms / 1e6 / sms / 1.965);
        CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        ms = timeit([&] { k_smem<<<sms * 2, 512, 65536>>>(out); });
        double bytes = (double)sms * 2 * 512 * (ITERS / 4) * 32 * 8;
        printf("SMEM LDS64+STS64: %.1f GB/s (%.1f B/clk/SM)\n", bytes / ms / 1e6, bytes / ms / 1e6 / sms / 1.965);
    }
    // memory
    size_t maxb = (size_t)2 << 30;
    float4* buf; CK(cudaMalloc(&buf, maxb)); CK(cudaMemset(buf, 0, maxb));
    size_t sizes_mb[] = {8, 16, 32, 48, 64, 80, 96, 128, 192, 2048};
    for (size_t mb : sizes_mb) {
        size_t n4 = (mb << 20) / 16;
        int reps = mb >= 1024 ? 2 : (int)(4096 / mb);
        float ms = timeit([&] { k_read<<<sms * 8, 256>>>(buf, n4, reps, out); });
        float msw = timeit([&] { k_write<<<sms * 8, 256>>>(buf, n4, reps, 1.f); });
        float msr = timeit([&] { k_rmw<<<sms * 8, 256>>>(buf, n4, reps); });
        printf("buf %5zu MB: read %.0f GB/s  write %.0f GB/s  rmw(r+w) %.0f GB/s\n", mb, (double)n4 * 16 * reps / ms / 1e6, (double)n4 * 16 * reps / msw / 1e6, 2.0 * n4 * 16 * reps / msr / 1e6);
    }
    for (int nimg : {2, 4, 8, 12}) {
        int M = 1024, reps = 64;
        float ms = timeit([&] { k_slab<<<sms * 2, 512>>>(buf, M, nimg, reps, out); });
        printf("slab read M=1024 nimg=%d (%d MB): %.0f GB/s\n", nimg, nimg * 8, (double)nimg * M * M * 8 * reps / ms / 1e6);
    }
    return 0;
}

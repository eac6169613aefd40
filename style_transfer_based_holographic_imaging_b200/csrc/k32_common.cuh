// k32_common.cuh -- constants, cp.async / prefetch helpers, table setup, H(z) phase, row load / emit stages
// Part of the FFT-size-1024 path; included by k32.cuh (which is included by asm_b200.cu).
#pragma once

namespace asmb {

constexpr int K32_L = 1024;
constexpr int K32_TW = 16 * 32;                       // half twiddle table entries (see fft_core.cuh, bfly HT)
#ifndef K32_ROW_WARPS_DEF
#define K32_ROW_WARPS_DEF 8
#endif
constexpr int K32_ROW_WARPS = K32_ROW_WARPS_DEF;      // rows in flight per CTA (register-landing row kernels)
constexpr int K32_ROW_CTAS = 16 / K32_ROW_WARPS;      // resident CTAs per SM the register-landing row kernels are compiled for
constexpr int K32_LP = RowLayout32::line_elems(K32_L);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// tw[e * 32 + Q] = W_{32 2^m}^{Q + 32 u}, e = off(m) + u, u < max(1, 2^{m-2}), off(1) = 0, off(m) = 2^{m-2}
// (half table: the other twiddles of a level are -i times these);  kappa table in natural column order.
__global__ void k32_setup(float2* tw, double* kzt, double s2, double inv_2pi_lambda) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int e = gtid; e < K32_TW; e += gsz) {
        const int ent = e / 32, Q = e % 32;
        int m = 1;
        while (m < 5 && ent >= (1 << (m - 1))) ++m;    // ent 0 -> m 1, 1 -> 2, 2..3 -> 3, 4..7 -> 4, 8..15 -> 5
        const int u = ent - (m == 1 ? 0 : (1 << (m - 2)));
        const int D = 32 << m, x = Q + 32 * u;
        float sn, cs;
        sincospif(2.0f * (float)x / (float)D, &sn, &cs);
        tw[e] = make_float2(cs, -sn);
    }
    constexpr int M = K32_L;
    for (int idx = gtid; idx < (M / 2 + 1) * M; idx += gsz) {
        const int ru = idx / M, v = idx % M;
        const int kv = v < M / 2 ? v : v - M;
        const double kk = (double)ru * ru + (double)kv * kv;
        const double arg = fma(-s2, kk, 1.0);
        kzt[idx] = (arg > 0.0 ? sqrt(arg) : 0.0) * inv_2pi_lambda;
    }
}

// multiply the column spectrum (v[i] = column frequency u = tl + 32 i of column c) by the transfer function:
// t = c_phase * kappa in fp64, reduced to [-1/2, 1/2] turns, sincos in fp32 (MUFU); DERIV: i kz H (grad_z)
template <bool DERIV, int CC>
__device__ __forceinline__ void k32_apply_h(float2 (&v)[32], const Params& p, const double* kz_s, int c, int tl, double cph) {
    constexpr int L = K32_L;
    const double MAGIC = 6755399441055744.0;                         // 1.5 * 2^52: round to nearest integer
    const double k2pl = 6.283185307179586 * p.lambda;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int u = tl + 32 * i;
        const int ru = u <= L / 2 ? u : L - u;
        const double kap = kz_s[ru * CC + c];
        const double tt = kap * cph;
        const double rr = tt - __dadd_rn(__dadd_rn(tt, MAGIC), -MAGIC);
        float sn, cn;
        __sincosf((float)rr * 6.283185307179586f, &sn, &cn);
        if constexpr (DERIV) v[i] = cmul_scaled(v[i], -sn, cn, (float)(kap * k2pl - p.kshift) * p.inv_m2);
        else v[i] = cmul_scaled(v[i], cn, sn, p.inv_m2);
    }
}

__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// cp.async.bulk.prefetch needs 16-byte aligned addresses and sizes: true when every row of the source qualifies
__device__ __forceinline__ bool k32_prefetch_ok(const Params& p) {
    if (p.N % 4 != 0) return false;
    uintptr_t a = 0;
    switch (p.in_mode) {
        case ASM_B200_IN_AMP_PHASE: case ASM_B200_IN_COT_FIELD: a = (uintptr_t)p.in0 | (uintptr_t)p.in1; break;
        case ASM_B200_IN_CONST_AMP_PHASE: a = (uintptr_t)p.in1; break;
        default: a = (uintptr_t)p.in0; break;
    }
    return (a & 15) == 0;
}

__device__ __forceinline__ void k32_prefetch_row(const Params& p, int plane, int y) {
    const size_t row = ((size_t)plane * p.N + y) * p.N;
    switch (p.in_mode) {
        case ASM_B200_IN_COMPLEX: l2_prefetch_bulk((const float2*)p.in0 + row, p.N * 8); break;
        case ASM_B200_IN_AMP_PHASE:
            l2_prefetch_bulk((const float*)p.in0 + row, p.N * 4);
            l2_prefetch_bulk((const float*)p.in1 + row, p.N * 4);
            break;
        case ASM_B200_IN_CONST_AMP_PHASE: l2_prefetch_bulk((const float*)p.in1 + row, p.N * 4); break;
        case ASM_B200_IN_COT_FIELD:
            l2_prefetch_bulk((const float*)p.in0 + row, p.N * 4);
            l2_prefetch_bulk((const float2*)p.in1 + row, p.N * 8);
            break;
        default: l2_prefetch_bulk((const float*)p.in0 + row, p.N * 4); break;
    }
}

template <int MODE>
__device__ __forceinline__ void load32(float2 (&v)[32], const Params& p, int plane, int y, int lane) {
    const size_t row = ((size_t)plane * p.N + y) * p.N;
    if (p.P == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = load_one<MODE>(p, row + lane + 32 * i);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            int x = lane + 32 * i - p.P;
            if (p.adj) {
                v[i] = (x >= 0 && x < p.N) ? load_one<MODE>(p, row + x) : make_float2(0.f, 0.f);
            } else {
                x = min(max(x, 0), p.N - 1);
                v[i] = load_one<MODE>(p, row + x);
            }
        }
    }
}

template <int MODE>
__device__ __forceinline__ float emit32(const float2 (&v)[32], const Params& p, int plane, int y, int lane, float2 fl, float2 fr) {
    float dot = 0.f;
    if (p.P == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) dot += emit_one<MODE>(p, plane, y, lane + 32 * i, v[i]);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int x = lane + 32 * i - p.P;
            if (x >= 0 && x < p.N) {
                float2 u = v[i];
                if (x == 0) { u.x += fl.x; u.y += fl.y; }
                if (x == p.N - 1) { u.x += fr.x; u.y += fr.y; }
                dot += emit_one<MODE>(p, plane, y, x, u);
            }
        }
    }
    return dot;
}

}  // namespace asmb

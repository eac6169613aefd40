// k32.cuh -- the FFT-size-1024 fast path: 32 points per thread, 1024 = 32 x 32.
//
// Shared-memory bandwidth (128 B/clk/SM) is the scarcest resource of this pipeline after the FP32 pipe, so this
// path makes ONE shared-memory exchange per 1-D FFT and moves everything else register <-> global directly:
//   k32_rows_fwd : one WARP per row.  lane l loads x[l + 32 i] (coalesced), radix-32 in registers, exchange
//                  through the warp's private 8.25 KB line (syncwarp only, no CTA barrier), radix-32 with table
//                  twiddles, and stores X[l + 32 i] straight from registers: register i of lane l holds frequency
//                  l + 32 i, so the workspace row is in NATURAL frequency order and the store is coalesced.
//   k32_cols     : 8 columns x 32 threads.  ws -> registers (64 B segments), radix-32, exchange, radix-32,
//                  x H(z) (kappa slab staged once per CTA with cp.async), inverse radix-32, exchange, radix-32,
//                  registers -> ws.
//   k32_rows_inv : mirror of k32_rows_fwd with the output stage fused.
// Included by asm_b200.cu (needs Params, load_one, emit_one, ...).
#pragma once

namespace asmb {

constexpr int K32_L = 1024;
constexpr int K32_TW = 31 * 32;                       // forward table entries
constexpr int K32_ROW_WARPS = 8;                      // rows in flight per CTA
#ifndef K32_ROW_CTAS_DEF
#define K32_ROW_CTAS_DEF 2
#endif
constexpr int K32_ROW_CTAS = K32_ROW_CTAS_DEF;             // resident CTAs per SM the LDG/STG row kernels are compiled for
#ifndef K32_NBUF_DEF
#define K32_NBUF_DEF 3
#endif
constexpr int K32_NBUF = K32_NBUF_DEF;                // line buffers per warp in the pipelined row kernels
constexpr int K32_LP = RowLayout32::line_elems(K32_L);
constexpr int K32_CC = 8;                             // columns per slab
constexpr int K32_SLAB_ROWS = ColLayout32<K32_CC>::rows(K32_L);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// tw32[e * 32 + Q] = W_{32 2^m}^{Q + 32 u}  (e = 2^{m-1}-1+u);  kappa table in natural column order
__global__ void k32_setup(float2* tw, double* kzt, int* ctl, int nctl, double s2, double inv_2pi_lambda) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    if (ctl) for (int i = gtid; i < nctl; i += gsz) ctl[i] = 0;
    for (int e = gtid; e < K32_TW; e += gsz) {
        const int ent = e / 32, Q = e % 32;
        int m = 1;
        while ((1 << m) - 1 <= ent) ++m;
        const int u = ent - ((1 << (m - 1)) - 1);
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
template <bool DERIV>
__device__ __forceinline__ void k32_apply_h(float2 (&v)[32], const Params& p, const double* kz_s, int c, int tl, double cph) {
    constexpr int L = K32_L, CC = K32_CC;
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

__device__ __forceinline__ void k32_prefetch_row(const Params& p, int plane, int y) {
    const size_t row = ((size_t)plane * p.N + y) * p.N;
    switch (p.in_mode) {
        case ASM_B200_IN_COMPLEX: l2_prefetch_bulk((const float2*)p.in0 + row, p.N * 8); break;
        case ASM_B200_IN_AMP_PHASE:
            l2_prefetch_bulk((const float*)p.in0 + row, p.N * 4);
            l2_prefetch_bulk((const float*)p.in1 + row, p.N * 4);
            break;
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

// forward row FFT of source row y of `plane` into workspace row `dst_row` (one warp; `line` is its private buffer)
__device__ __forceinline__ void k32_row_fwd(const Params& p, float2* line, const float2* tw, int lane, int plane, int y,
                                            float2* dst_row) {
        float2 v[32];
        if (p.dbg & 1024) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = make_float2((float)(lane + i), (float)(y - i));
        } else
        switch (p.in_mode) {
            case ASM_B200_IN_COMPLEX: load32<ASM_B200_IN_COMPLEX>(v, p, plane, y, lane); break;
            case ASM_B200_IN_AMP_PHASE: load32<ASM_B200_IN_AMP_PHASE>(v, p, plane, y, lane); break;
            case ASM_B200_IN_SQRT_REAL: load32<ASM_B200_IN_SQRT_REAL>(v, p, plane, y, lane); break;
            case ASM_B200_IN_COT_FIELD: load32<ASM_B200_IN_COT_FIELD>(v, p, plane, y, lane); break;
            default: load32<ASM_B200_IN_REAL>(v, p, plane, y, lane); break;
        }
        if (!(p.dbg & 256)) {
        fwd32_first(v);                                              // digit of bits 5..9 (positions lane + 32 i)
        sts16<RowLayout32, 5>(v, line + lane);
        __syncwarp();
        lds16<RowLayout32, 0>(v, line + 33 * lane);                  // positions 32 lane + i
        fwd32_table(v, tw + lane);
        __syncwarp();                                                // line is rewritten by the next row
        }
        float2* dst = dst_row + lane;                                // frequency lane + 32 i
        if (p.dbg & 512) {
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += v[i].x + v[i].y;
            if (acc == 1.2345e33f) dst[0] = v[0];
        } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) __stcg(dst + 32 * i, v[i]);
        }
}

__global__ void __launch_bounds__(32 * K32_ROW_WARPS, K32_ROW_CTAS) k32_rows_fwd(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* lines = reinterpret_cast<float2*>(smem_raw);             // [K32_ROW_WARPS][K32_LP]
    float2* tw = lines + K32_ROW_WARPS * K32_LP;                     // [31][32]
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    __syncthreads();
    float2* line = lines + w * K32_LP;
    const bool prefetch = !(p.dbg & 32) && p.N % 4 == 0;
    for (int gline = blockIdx.x * K32_ROW_WARPS + w; gline < nlines; gline += gridDim.x * K32_ROW_WARPS) {
        const int img = gline / p.N, y = gline % p.N;
        // pull this warp's NEXT source row from HBM into L2 while this one is transformed
        const int nxt = gline + gridDim.x * K32_ROW_WARPS;
        if (prefetch && lane == 0 && nxt < nlines) k32_prefetch_row(p, plane0 + nxt / p.N, nxt % p.N);
        const int wrow = (p.dbg & 2048) ? gline % (4 * p.N) : gline;   // timing experiment: keep the writes inside 32 MB
        k32_row_fwd(p, line, tw, lane, plane0 + img, y, p.ws + (size_t)wrow * K32_L);
    }
}

// inverse row FFT of workspace row `src_row` + output stage for row y of `plane` (one warp)
__device__ __forceinline__ void k32_row_inv(const Params& p, float2* line, const float2* tw, int lane, int plane, int y,
                                            const float2* src_row) {
        const bool folding = p.adj && p.P > 0;
        float2 v[32];
        const float2* src = src_row + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __ldcg(src + 32 * i);   // frequency lane + 32 i = position 32 lane + i
        if (!(p.dbg & 16)) {
            // the intermediate row is dead now: drop its (dirty) L2 lines instead of letting them be written back to HBM
            // (the slot is completely rewritten by the next forward row pass before anything reads it again)
            float sink = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) sink += v[i].x;               // order the discard after the loads have returned
            if (sink != 1.2345e-33f) {
                asm volatile("discard.global.L2 [%0], 128;" ::"l"((const char*)src_row + (size_t)lane * 128) : "memory");
                asm volatile("discard.global.L2 [%0], 128;" ::"l"((const char*)src_row + (size_t)(lane + 32) * 128) : "memory");
            }
        }
        inv32_first(v);
        sts16<RowLayout32, 0>(v, line + 33 * lane);
        __syncwarp();
        lds16<RowLayout32, 5>(v, line + lane);
        inv32_table(v, tw + lane);                                   // v[i] = natural position lane + 32 i
        __syncwarp();
        float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
        if (folding) {
            // adjoint of replicate padding: fold columns [0,P) onto 0 and [P+N, M) onto N-1 (warp reduction)
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int pos = lane + 32 * i;
                if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                if (pos >= p.P + p.N) { fr.x += v[i].x; fr.y += v[i].y; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fl.x += __shfl_xor_sync(0xffffffffu, fl.x, o); fl.y += __shfl_xor_sync(0xffffffffu, fl.y, o);
                fr.x += __shfl_xor_sync(0xffffffffu, fr.x, o); fr.y += __shfl_xor_sync(0xffffffffu, fr.y, o);
            }
        }
        float dot = 0.f;
        switch (p.out_mode) {
            case ASM_B200_OUT_COMPLEX: emit32<ASM_B200_OUT_COMPLEX>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_INTENSITY: emit32<ASM_B200_OUT_INTENSITY>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_ABS_ANGLE: emit32<ASM_B200_OUT_ABS_ANGLE>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_REIM_CAT: emit32<ASM_B200_OUT_REIM_CAT>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_ABSANG_CAT: emit32<ASM_B200_OUT_ABSANG_CAT>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_GRAD_AP: emit32<ASM_B200_OUT_GRAD_AP>(v, p, plane, y, lane, fl, fr); break;
            default: dot = emit32<OUT_DOT>(v, p, plane, y, lane, fl, fr); break;
        }
        if (p.out_mode == OUT_DOT) {   // a warp owns one row of one sample
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const double K = p.z_f64 ? 6.283185307179586 : (double)6.2831854820251465f;
            if (lane == 0) atomicAdd((double*)p.out0 + plane / p.C, (double)dot * K * p.inv_lambda);
        }
}

__global__ void __launch_bounds__(32 * K32_ROW_WARPS, K32_ROW_CTAS) k32_rows_inv(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* lines = reinterpret_cast<float2*>(smem_raw);
    float2* tw = lines + K32_ROW_WARPS * K32_LP;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    __syncthreads();
    float2* line = lines + w * K32_LP;
    for (int gline = blockIdx.x * K32_ROW_WARPS + w; gline < nlines; gline += gridDim.x * K32_ROW_WARPS) {
        const int img = gline / p.N, y = gline % p.N;
        k32_row_inv(p, line, tw, lane, plane0 + img, y, p.ws + ((size_t)img * p.N + y) * K32_L);
    }
}


// One slab (8 columns) of the image whose N workspace rows start at `img_ws`: 256 threads, c = t % 8, tl = t / 8.
// smem: slab | kappa slab | (tw, fold owned by the caller).  kz_loaded: the kappa slab of this column block is
// already resident in kz_s (persistent callers); otherwise it is staged here with cp.async.
__device__ __forceinline__ void k32_col_slab(const Params& p, float2* slab, double* kz_s, const float2* tw, float2* fold,
                                             int plane, int slab_i, float2* img_ws, bool kz_loaded) {
    constexpr int L = K32_L, CC = K32_CC;
    using LAY = ColLayout32<CC>;
    const int t = threadIdx.x, c = t % CC, tl = t / CC;
    const int col0 = slab_i * CC;

    // stage the kappa slab (64 B per frequency row) asynchronously; it is consumed after the first barrier
    if (!kz_loaded) {
        for (int j = t; j < (L / 2 + 1) * 4; j += 32 * CC) {
            const int ru = j >> 2, q = j & 3;
            cp_async16(kz_s + ru * CC + 2 * q, p.kzt + (size_t)ru * L + col0 + 2 * q);
        }
    }
    if (t < 2 * CC) fold[t] = make_float2(0.f, 0.f);

    // ---- load: rows tl + 32 i of column col0 + c (padding rows by clamp / zero) ----
    float2 v[32];
    const float2* src = img_ws + col0 + c;
    if (p.P == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __ldcg(src + (size_t)(tl + 32 * i) * L);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            int r = tl + 32 * i - p.P;
            if (p.adj) v[i] = (r >= 0 && r < p.N) ? __ldcg(src + (size_t)r * L) : make_float2(0.f, 0.f);
            else { r = min(max(r, 0), p.N - 1); v[i] = __ldcg(src + (size_t)r * L); }
        }
    }
    const int b = plane / p.C;
    double cph;                                                      // phase constant c (ASM.py:29)
    if (p.z_f64) cph = 6.283185307179586 * __ldg((const double*)p.z + b);
    else cph = (double)__fmul_rn(6.2831854820251465f, __ldg((const float*)p.z + b));
    if (p.h_mode == H_CONJ) cph = -cph;

    float2* col = slab + c;
    // ---- forward column FFT ----
    fwd32_first(v);
    sts16<LAY, 5>(v, col + tl * CC);                                 // rows tl + 32 i  -> padded rows tl + 33 i
    cp_async_wait_all();
    __syncthreads();
    lds16<LAY, 0>(v, col + 33 * tl * CC);                            // rows 32 tl + i
    fwd32_table(v, tw + tl);                                         // v[i] = column frequency u = tl + 32 i

    // ---- transfer function ----
    {
            if (p.h_mode == H_DERIV) k32_apply_h<true>(v, p, kz_s, c, tl, cph);
            else k32_apply_h<false>(v, p, kz_s, c, tl, cph);
        }

    // ---- inverse column FFT ----
    inv32_first(v);
    sts16<LAY, 0>(v, col + 33 * tl * CC);
    __syncthreads();
    lds16<LAY, 5>(v, col + tl * CC);
    inv32_table(v, tw + tl);                                         // v[i] = row tl + 32 i, natural order

    // ---- store rows [P, P+N) (crop); adjoint: fold the padding rows onto rows P and P+N-1 first ----
    float2* dst = img_ws + col0 + c;
    if (p.P == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) __stcg(dst + (size_t)(tl + 32 * i) * L, v[i]);
    } else {
        float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
        if (p.adj) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int pos = tl + 32 * i;
                if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                if (pos >= p.P + p.N) { fr.x += v[i].x; fr.y += v[i].y; }
            }
            atomicAdd(&fold[c].x, fl.x); atomicAdd(&fold[c].y, fl.y);
            atomicAdd(&fold[CC + c].x, fr.x); atomicAdd(&fold[CC + c].y, fr.y);
            __syncthreads();
            fl = fold[c]; fr = fold[CC + c];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int r = tl + 32 * i - p.P;
            if (r >= 0 && r < p.N) {
                float2 u = v[i];
                if (r == 0) { u.x += fl.x; u.y += fl.y; }
                if (r == p.N - 1) { u.x += fr.x; u.y += fr.y; }
                __stcg(dst + (size_t)r * L, u);
            }
        }
    }
}

__global__ void __launch_bounds__(32 * K32_CC, 2) k32_cols(const Params p, int plane0, int nimg) {
    constexpr int L = K32_L, CC = K32_CC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* slab = reinterpret_cast<float2*>(smem_raw);              // [K32_SLAB_ROWS][CC]
    double* kz_s = reinterpret_cast<double*>(slab + K32_SLAB_ROWS * CC);  // [L/2+1][CC]
    float2* tw = reinterpret_cast<float2*>(kz_s + (L / 2 + 1) * CC); // [31][32]
    float2* fold = tw + K32_TW;                                      // [2][CC]
    constexpr int nslab = L / CC;
    for (int i = threadIdx.x; i < K32_TW; i += 32 * CC) tw[i] = __ldg(p.tw + i);
    // persistent: slab-major order so that a CTA mostly keeps its kappa slab resident across samples
    int kz_slab = -1;
    for (int wi = blockIdx.x; wi < nimg * nslab; wi += gridDim.x) {
        const int img = wi / nslab, slab_i = wi % nslab;   // image-major: neighbouring CTAs read neighbouring 64 B segments
        __syncthreads();
        k32_col_slab(p, slab, kz_s, tw, fold, plane0 + img, slab_i, p.ws + (size_t)img * p.N * L, kz_slab == slab_i);
        kz_slab = slab_i;
    }
}

// ---------------------------------------------------------------------------------------------------
// Pipelined row kernels (default for complex64 / amplitude+phase input): ONE persistent CTA of 8 warps per SM.
// Every warp owns two 8.25 KB line buffers: while it transforms the row in one of them, cp.async lands its next
// row in the other (no registers tied up by loads in flight, no warp waiting on DRAM / L2).  The buffer that
// held the raw row doubles as the exchange buffer once the row is in registers.  No CTA barrier in the loop.
// ---------------------------------------------------------------------------------------------------
// stage `bytes` (multiple of 16) from gmem to smem with this warp's 32 lanes
__device__ __forceinline__ void warp_stage(void* dst, const void* src, int bytes, int lane) {
    for (int o = lane * 16; o < bytes; o += 32 * 16) cp_async16((char*)dst + o, (const char*)src + o);
}

__global__ void __launch_bounds__(32 * K32_ROW_WARPS, 1) k32_rows_fwd_pipe(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* lines = reinterpret_cast<float2*>(smem_raw);             // [K32_ROW_WARPS][K32_NBUF][K32_LP]
    float2* tw = lines + K32_ROW_WARPS * K32_NBUF * K32_LP;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    __syncthreads();
    float2* base = lines + (size_t)w * K32_NBUF * K32_LP;
    const bool ap = p.in_mode == ASM_B200_IN_AMP_PHASE;
    const int stride = gridDim.x * K32_ROW_WARPS;
    auto stage = [&](float2* dst, int gline) {
        const int img = gline / p.N, y = gline % p.N;
        const size_t row = ((size_t)(plane0 + img) * p.N + y) * p.N;
        if (ap) {
            warp_stage(dst, (const float*)p.in0 + row, p.N * 4, lane);
            warp_stage((float*)dst + p.N, (const float*)p.in1 + row, p.N * 4, lane);
        } else {
            warp_stage(dst, (const float2*)p.in0 + row, p.N * 8, lane);
        }
    };
    int gline = blockIdx.x * K32_ROW_WARPS + w;
#pragma unroll
    for (int k = 0; k < K32_NBUF - 1; ++k) {                         // prologue: NBUF-1 rows in flight
        if (gline + k * stride < nlines) stage(base + k * K32_LP, gline + k * stride);
        cp_async_commit();
    }
    for (int it = 0; gline < nlines; gline += stride, ++it) {
        float2* cur = base + (it % K32_NBUF) * K32_LP;
        if (gline + (K32_NBUF - 1) * stride < nlines) stage(base + ((it + K32_NBUF - 1) % K32_NBUF) * K32_LP, gline + (K32_NBUF - 1) * stride);
        cp_async_commit();
        cp_async_wait<K32_NBUF - 1>();                               // the current row has landed
        __syncwarp();
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            int x = lane + 32 * i - p.P;
            bool in = true;
            if (p.P != 0) { in = !p.adj || (x >= 0 && x < p.N); x = min(max(x, 0), p.N - 1); }
            if (ap) {
                const float a = ((const float*)cur)[x];
                const float ph = ((const float*)cur)[p.N + x] * p.in_scale;
                float sn, cs;
                sincos_full(ph, &sn, &cs);
                v[i] = in ? make_float2(a * cs, a * sn) : make_float2(0.f, 0.f);
            } else {
                v[i] = in ? cur[x] : make_float2(0.f, 0.f);
            }
        }
        __syncwarp();                                                // raw row consumed: `cur` becomes the exchange line
        fwd32_first(v);
        sts16<RowLayout32, 5>(v, cur + lane);
        __syncwarp();
        lds16<RowLayout32, 0>(v, cur + 33 * lane);
        fwd32_table(v, tw + lane);
        __syncwarp();
        const int img = gline / p.N, y = gline % p.N;
        float2* dst = p.ws + ((size_t)img * p.N + y) * K32_L + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) __stcg(dst + 32 * i, v[i]);
    }
    cp_async_wait<0>();
}

__global__ void __launch_bounds__(32 * K32_ROW_WARPS, 1) k32_rows_inv_pipe(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* lines = reinterpret_cast<float2*>(smem_raw);
    float2* tw = lines + K32_ROW_WARPS * K32_NBUF * K32_LP;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    __syncthreads();
    float2* base = lines + (size_t)w * K32_NBUF * K32_LP;
    const int stride = gridDim.x * K32_ROW_WARPS;
    const bool folding = p.adj && p.P > 0;
    int gline = blockIdx.x * K32_ROW_WARPS + w;
#pragma unroll
    for (int k = 0; k < K32_NBUF - 1; ++k) {
        if (gline + k * stride < nlines) warp_stage(base + k * K32_LP, p.ws + (size_t)(gline + k * stride) * K32_L, K32_L * 8, lane);
        cp_async_commit();
    }
    for (int it = 0; gline < nlines; gline += stride, ++it) {
        float2* cur = base + (it % K32_NBUF) * K32_LP;
        if (gline + (K32_NBUF - 1) * stride < nlines)
            warp_stage(base + ((it + K32_NBUF - 1) % K32_NBUF) * K32_LP, p.ws + (size_t)(gline + (K32_NBUF - 1) * stride) * K32_L, K32_L * 8, lane);
        cp_async_commit();
        cp_async_wait<K32_NBUF - 1>();
        __syncwarp();
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cur[lane + 32 * i];     // frequency lane + 32 i = position 32 lane + i
        __syncwarp();
        inv32_first(v);
        sts16<RowLayout32, 0>(v, cur + 33 * lane);
        __syncwarp();
        lds16<RowLayout32, 5>(v, cur + lane);
        inv32_table(v, tw + lane);                                   // v[i] = natural position lane + 32 i
        __syncwarp();
        const int img = gline / p.N, y = gline % p.N, plane = plane0 + img;
        float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
        if (folding) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int pos = lane + 32 * i;
                if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                if (pos >= p.P + p.N) { fr.x += v[i].x; fr.y += v[i].y; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fl.x += __shfl_xor_sync(0xffffffffu, fl.x, o); fl.y += __shfl_xor_sync(0xffffffffu, fl.y, o);
                fr.x += __shfl_xor_sync(0xffffffffu, fr.x, o); fr.y += __shfl_xor_sync(0xffffffffu, fr.y, o);
            }
        }
        float dot = 0.f;
        switch (p.out_mode) {
            case ASM_B200_OUT_COMPLEX: emit32<ASM_B200_OUT_COMPLEX>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_INTENSITY: emit32<ASM_B200_OUT_INTENSITY>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_ABS_ANGLE: emit32<ASM_B200_OUT_ABS_ANGLE>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_REIM_CAT: emit32<ASM_B200_OUT_REIM_CAT>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_ABSANG_CAT: emit32<ASM_B200_OUT_ABSANG_CAT>(v, p, plane, y, lane, fl, fr); break;
            case ASM_B200_OUT_GRAD_AP: emit32<ASM_B200_OUT_GRAD_AP>(v, p, plane, y, lane, fl, fr); break;
            default: dot = emit32<OUT_DOT>(v, p, plane, y, lane, fl, fr); break;
        }
        if (p.out_mode == OUT_DOT) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const double K = p.z_f64 ? 6.283185307179586 : (double)6.2831854820251465f;
            if (lane == 0) atomicAdd((double*)p.out0 + plane / p.C, (double)dot * K * p.inv_lambda);
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------
// Pipelined column kernel (default): ONE persistent CTA per SM.  The next slab is copied global -> shared with
// cp.async (no registers, no waiting warps) while the current one is transformed, and results leave straight
// from registers; the kappa slab of the next item is fetched during the current inverse transform.
//   smem: raw [N rows][8] (dense landing zone) | exchange slab [1056][8] (padded) | kappa [513][8] | tw | fold
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ void k32_stage_raw(float2* raw, const float2* img_ws, int col0, int nrows) {
    for (int j = threadIdx.x; j < nrows * 4; j += 32 * K32_CC) {
        const int r = j >> 2, q = j & 3;
        cp_async16(raw + r * K32_CC + 2 * q, img_ws + (size_t)r * K32_L + col0 + 2 * q);
    }
}
__device__ __forceinline__ void k32_stage_kz(double* kz_s, const double* kzt, int col0) {
    for (int j = threadIdx.x; j < (K32_L / 2 + 1) * 4; j += 32 * K32_CC) {
        const int ru = j >> 2, q = j & 3;
        cp_async16(kz_s + ru * K32_CC + 2 * q, kzt + (size_t)ru * K32_L + col0 + 2 * q);
    }
}

// SHARED = false: separate landing zone, 1 CTA/SM, prefetch right after the raw slab is consumed.
// SHARED = true : the landing zone IS the exchange slab (dense rows in its first 64 KB), 2 CTAs/SM, the next slab
//                 is prefetched once the last exchange read of the current item is done.
// Items first, first + step, ... < total (item = img * 128 + slab; image img lives in workspace slot img % ring).
// The twiddle table must already be in `tw`; every thread of the CTA calls this.
template <bool SHARED>
__device__ __forceinline__ void k32_cols_items(const Params& p, float2* raw, float2* slab, double* kz_s, const float2* tw, float2* fold,
                                               int plane0, int first, int total, int step, int ring) {
    constexpr int L = K32_L, CC = K32_CC, nslab = L / CC;
    using LAY = ColLayout32<CC>;
    const int t = threadIdx.x, c = t % CC, tl = t / CC;
    int wi = first;
    if (wi < total) {                                                // prologue: first slab + its kappa
        k32_stage_raw(raw, p.ws + (size_t)((wi / nslab) % ring) * p.N * L, (wi % nslab) * CC, p.N);
        cp_async_commit();
        k32_stage_kz(kz_s, p.kzt, (wi % nslab) * CC);
        cp_async_commit();
    }
    float2* col = slab + c;
    for (; wi < total; wi += step) {
        const int img = wi / nslab, slab_i = wi % nslab, plane = plane0 + img;
        const int col0 = slab_i * CC;
        const int nxt = wi + step;
        float2* img_ws = p.ws + (size_t)(img % ring) * p.N * L;
        if (t < 2 * CC) fold[t] = make_float2(0.f, 0.f);
        if (SHARED) cp_async_wait<0>(); else cp_async_wait<1>();     // this item's raw slab has landed
        __syncthreads();
        // ---- registers <- raw rows tl + 32 i (padding rows by clamp / zero) ----
        float2 v[32];
        if (p.P == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = raw[(tl + 32 * i) * CC + c];
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                int r = tl + 32 * i - p.P;
                if (p.adj) v[i] = (r >= 0 && r < p.N) ? raw[r * CC + c] : make_float2(0.f, 0.f);
                else { r = min(max(r, 0), p.N - 1); v[i] = raw[r * CC + c]; }
            }
        }
        __syncthreads();                                             // raw is free
        if (!SHARED) {                                               // ... prefetch the next slab into it right away
            if (nxt < total && !(p.dbg & 2)) k32_stage_raw(raw, p.ws + (size_t)((nxt / nslab) % ring) * p.N * L, (nxt % nslab) * CC, p.N);
            cp_async_commit();
        }

        const int b = plane / p.C;
        double cph;                                                  // phase constant c (ASM.py:29)
        if (p.z_f64) cph = 6.283185307179586 * __ldg((const double*)p.z + b);
        else cph = (double)__fmul_rn(6.2831854820251465f, __ldg((const float*)p.z + b));
        if (p.h_mode == H_CONJ) cph = -cph;

        // ---- forward column FFT ----
        fwd32_first(v);
        sts16<LAY, 5>(v, col + tl * CC);
        if (!SHARED) cp_async_wait<1>();                             // kappa of this item (committed before the raw prefetch)
        __syncthreads();
        lds16<LAY, 0>(v, col + 33 * tl * CC);
        fwd32_table(v, tw + tl);
        // ---- transfer function ----
        if (!(p.dbg & 8)) {
            if (p.h_mode == H_DERIV) k32_apply_h<true>(v, p, kz_s, c, tl, cph);
            else k32_apply_h<false>(v, p, kz_s, c, tl, cph);
        }
        // ---- inverse column FFT ----
        inv32_first(v);
        sts16<LAY, 0>(v, col + 33 * tl * CC);
        __syncthreads();                                             // every thread is done with kz_s too
        if (nxt < total && (nxt % nslab) != slab_i && !(p.dbg & 4)) k32_stage_kz(kz_s, p.kzt, (nxt % nslab) * CC);
        cp_async_commit();
        lds16<LAY, 5>(v, col + tl * CC);
        if (SHARED) {                                                // the slab is dead from here on: land the next one in it
            __syncthreads();
            if (nxt < total && !(p.dbg & 2)) k32_stage_raw(raw, p.ws + (size_t)((nxt / nslab) % ring) * p.N * L, (nxt % nslab) * CC, p.N);
            cp_async_commit();
        }
        inv32_table(v, tw + tl);
        // ---- store rows [P, P+N) (crop); adjoint: fold the padding rows onto rows P and P+N-1 first ----
        float2* dst = img_ws + col0 + c;
        if (p.dbg & 1) {   // timing experiment: no stores (keep the values alive)
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += v[i].x + v[i].y;
            if (acc == 1.2345e33f) dst[0] = v[0];
        } else if (p.P == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(dst + (size_t)(tl + 32 * i) * L, v[i]);
        } else {
            float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
            if (p.adj) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pos = tl + 32 * i;
                    if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= p.P + p.N) { fr.x += v[i].x; fr.y += v[i].y; }
                }
                atomicAdd(&fold[c].x, fl.x); atomicAdd(&fold[c].y, fl.y);
                atomicAdd(&fold[CC + c].x, fr.x); atomicAdd(&fold[CC + c].y, fr.y);
                __syncthreads();
                fl = fold[c]; fr = fold[CC + c];
                __syncthreads();                                     // fold is re-zeroed at the top of the next item
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int r = tl + 32 * i - p.P;
                if (r >= 0 && r < p.N) {
                    float2 u = v[i];
                    if (r == 0) { u.x += fl.x; u.y += fl.y; }
                    if (r == p.N - 1) { u.x += fr.x; u.y += fr.y; }
                    __stcg(dst + (size_t)r * L, u);
                }
            }
        }
    }
    cp_async_wait<0>();
}

template <bool SHARED>
__global__ void __launch_bounds__(32 * K32_CC, SHARED ? 2 : 1) k32_cols_pipe(const Params p, int plane0, int nimg) {
    constexpr int L = K32_L, CC = K32_CC, nslab = L / CC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* raw = reinterpret_cast<float2*>(smem_raw);               // [L][CC] dense (N rows used)
    float2* slab = SHARED ? raw : raw + L * CC;                      // [K32_SLAB_ROWS][CC]
    double* kz_s = reinterpret_cast<double*>(slab + K32_SLAB_ROWS * CC);
    float2* tw = reinterpret_cast<float2*>(kz_s + (L / 2 + 1) * CC);
    float2* fold = tw + K32_TW;
    for (int i = threadIdx.x; i < K32_TW; i += 32 * CC) tw[i] = __ldg(p.tw + i);
    k32_cols_items<SHARED>(p, raw, slab, kz_s, tw, fold, plane0, blockIdx.x, nimg * nslab, gridDim.x, 0x7fffffff);
}

// ---------------------------------------------------------------------------------------------------
// Persistent dataflow kernel: ONE launch per call, two kinds of resident workers.
//   row workers (first half of the grid): every WARP is independent -- it pulls row tickets
//        step s:  [forward rows of image s] [inverse rows of image s-2]      (K32_RPT rows per ticket)
//     and never meets a CTA barrier; these warps stream HBM <-> L2 and fill the issue slots that the
//     barrier-synchronised column worker on the same SM leaves idle.
//   column workers (second half): one CTA per slab ticket (image-major), k32_col_slab.
// Image b lives in ring slot b % R of the L2-resident workspace.  Dependencies are per-image counters
//   done1[b] forward rows written, done2[b] slabs done, done3[b] inverse rows consumed (slot may be reused);
// each chain of waits strictly decreases in b or moves to an earlier ticket of an in-order queue, so it
// terminates whatever the residency.  ctl[0] row ticket, ctl[1] column ticket, ctl[32...] the counters.
// ---------------------------------------------------------------------------------------------------
constexpr int K32_RPT = 4;   // rows per row ticket

__device__ __forceinline__ int ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256, 2) k32_mega(const Params p, int* ctl, int R, int nowait) {
    constexpr int L = K32_L, CC = K32_CC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);               // 8 row lines, or one column slab
    double* kz_s = reinterpret_cast<double*>(buf + K32_SLAB_ROWS * CC);
    float2* tw = reinterpret_cast<float2*>(kz_s + (L / 2 + 1) * CC);
    float2* fold = tw + K32_TW;
    int* s_tick = reinterpret_cast<int*>(fold + 2 * CC);
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    int* done1 = ctl + 32;
    int* done2 = done1 + p.planes;
    int* done3 = done2 + p.planes;
    constexpr int n2 = L / CC;                                       // column slabs per image

    for (int i = t; i < K32_TW; i += 256) tw[i] = __ldg(p.tw + i);
    __syncthreads();

    if (blockIdx.x < gridDim.x / 2) {
        // ------------------------------ row worker: warps are independent ------------------------------
        const int n1 = p.N / K32_RPT;                                // tickets per image and direction
        const int total = (p.planes + 2) * 2 * n1;
        float2* line = buf + w * K32_LP;
        for (;;) {
            int tk = 0;
            if (lane == 0) tk = atomicAdd(ctl, 1);
            tk = __shfl_sync(0xffffffffu, tk, 0);
            if (tk >= total) break;
            const int s = tk / (2 * n1), r = tk - s * 2 * n1;
            const bool fwd = r < n1;
            const int b = fwd ? s : s - 2;
            if (b < 0 || b >= p.planes) continue;
            const int y0 = (fwd ? r : r - n1) * K32_RPT;
            if (lane == 0 && !nowait) {
                if (fwd) { if (b >= R) while (ld_acquire(done3 + (b - R)) < p.N) __nanosleep(100); }
                else while (ld_acquire(done2 + b) < n2) __nanosleep(100);
            }
            __syncwarp();
            float2* img_ws = p.ws + (size_t)(b % R) * p.N * L;
            if (fwd && p.in_mode == ASM_B200_IN_COMPLEX && p.P == 0) {
                // pull the rows of the NEXT ticket of this warp's neighbourhood from HBM into L2 while this one is
                // transformed (the loads below then cost an L2 hit instead of a DRAM round trip)
                const int ty = y0 + 8 * K32_RPT;                    // ~8 tickets ahead in the same image
                if (ty + K32_RPT <= p.N) {
                    const char* pf = (const char*)((const float2*)p.in0 + ((size_t)b * p.N + ty) * p.N);
#pragma unroll
                    for (int q = 0; q < K32_RPT * 2; ++q)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (size_t)(q * 32 + lane) * 128));
                }
            }
#pragma unroll 1
            for (int j = 0; j < K32_RPT; ++j) {
                const int y = y0 + j;
                if (fwd) k32_row_fwd(p, line, tw, lane, b, y, img_ws + (size_t)y * L);
                else k32_row_inv(p, line, tw, lane, b, y, img_ws + (size_t)y * L);
            }
            __syncwarp();
            if (lane == 0) { __threadfence(); atomicAdd((fwd ? done1 : done3) + b, K32_RPT); }
        }
    } else {
        // ------------------------------ column worker: one slab per ticket ------------------------------
        const int total = p.planes * n2;
        int kz_slab = -1;                                            // kappa slab currently resident in kz_s
        for (;;) {
            if (t == 0) s_tick[0] = atomicAdd(ctl + 1, 1);
            __syncthreads();
            const int tk = s_tick[0];
            if (tk >= total) break;
            const int b = tk / n2, item = tk - b * n2;
            if (t == 0 && !nowait) while (ld_acquire(done1 + b) < p.N) __nanosleep(100);
            __syncthreads();
            k32_col_slab(p, buf, kz_s, tw, fold, b, item, p.ws + (size_t)(b % R) * p.N * L, kz_slab == item);
            kz_slab = item;
            __syncthreads();
            if (t == 0) { __threadfence(); atomicAdd(done2 + b, 1); }
        }
    }
}

}  // namespace asmb

namespace asmb {

// ---------------------------------------------------------------------------------------------------
// k32_flow: the whole call as ONE persistent launch with UNIFORM workers and a tight L2-resident ring.
// Work items ("tickets"), per image b:
//   F(b, g): forward row FFTs of rows [rpt g, rpt g + rpt)               (8 independent warps x rpt/8 rows)
//   C(b, j): column slabs [cq j, cq j + cq) (FFT . H(z) . IFFT in place)  (the CTA as 8 columns x 32 threads)
//   I(b, g): inverse row FFTs + output stage of rows [rpt g, rpt g + rpt)
// Images in the window [lo, lo + R) are active (lo = oldest image whose output is incomplete); image b lives in
// ring slot b % R.  Every resident CTA (2 per SM) repeatedly lets its warp 0 look at the window (one lane per image,
// six counters each, one memory round trip) and claims, by atomicAdd on a per-image per-pass counter, a ticket of
// the oldest image with READY work, trying I, then C, then F:  I(b) is ready when all slabs of b are done, C(b) when
// all its forward rows are written, F(b) as soon as b is inside the window.  Only work whose dependencies are
// COMPLETE is ever claimed (an overshooting atomicAdd yields no ticket, never a wrong one), so no worker waits while
// holding a ticket: the schedule is work conserving and cannot deadlock whatever the residency, and R ~ 4-6 slots
// suffice -- the intermediate never leaves L2.  At any time the resident tickets are a mix of HBM-reading,
// compute-bound and HBM-writing work.
// ctl: [32 + k planes + b], k = 0..5: claimF, claimC, claimI, done1 (rows written), done2 (slabs), done3 (rows out).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k32_flow(const Params p, int* ctl, int R, int rpt, int cq) {
    constexpr int L = K32_L, CC = K32_CC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);               // 8 row lines, or one column slab
    double* kz_s = reinterpret_cast<double*>(buf + K32_SLAB_ROWS * CC);
    float2* tw = reinterpret_cast<float2*>(kz_s + (L / 2 + 1) * CC);
    float2* fold = tw + K32_TW;
    int* s_tick = reinterpret_cast<int*>(fold + 2 * CC);
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    int* claimF = ctl + 32;
    int* claimC = claimF + p.planes;
    int* claimI = claimC + p.planes;
    int* done1 = claimI + p.planes;
    int* done2 = done1 + p.planes;
    int* done3 = done2 + p.planes;
    constexpr int nslab = L / CC;                                    // column slabs per image
    const int nF = p.N / rpt;                                        // row tickets per image and direction
    const int nC = nslab / cq;                                       // column tickets per image
    const bool prefetch = !(p.dbg & 32);
    const int rpw = rpt / 8;                                         // rows per warp and ticket

    for (int i = t; i < K32_TW; i += 256) tw[i] = __ldg(p.tw + i);
    float2* line = buf + w * K32_LP;
    int lo = 0;                                                      // warp 0: oldest image not known to be complete

    for (;;) {
        __syncthreads();                                             // previous ticket done with smem (and s_tick)
        if (w == 0) {
            int kind = 3, img = 0, tk = 0;
            while (lo < p.planes) {
                const int b = lo + lane;
                const bool act = lane < R && b < p.planes;
                int cF = nF, cC = nC, cI = nF, d1 = 0, d2 = 0, d3 = 0;
                if (act) {   // six independent relaxed loads (one round trip); the acquire fence follows the claim
                    cF = ld_relaxed(claimF + b); cC = ld_relaxed(claimC + b); cI = ld_relaxed(claimI + b);
                    d1 = ld_relaxed(done1 + b); d2 = ld_relaxed(done2 + b); d3 = ld_relaxed(done3 + b);
                }
                // slide the window over the leading complete images
                const unsigned incomplete = __ballot_sync(0xffffffffu, !act || d3 < p.N);
                const int adv = __ffs(incomplete) - 1;               // lanes [0, adv) hold complete images
                if (adv > 0) { lo += adv; continue; }
                const unsigned rI = __ballot_sync(0xffffffffu, act && d2 >= nslab && cI < nF);
                const unsigned rC = __ballot_sync(0xffffffffu, act && d1 >= p.N && cC < nC);
                const unsigned rF = __ballot_sync(0xffffffffu, act && cF < nF);
                int k = -1, sel = 0;
                if (rI) { k = 2; sel = __ffs(rI) - 1; }
                else if (rC) { k = 1; sel = __ffs(rC) - 1; }
                else if (rF) { k = 0; sel = __ffs(rF) - 1; }
                if (k < 0) { __nanosleep(256); continue; }
                int got = 0;
                if (lane == 0) {
                    int* cnt = (k == 2 ? claimI : k == 1 ? claimC : claimF) + lo + sel;
                    got = atomicAdd(cnt, 1);
                }
                got = __shfl_sync(0xffffffffu, got, 0);
                if (got < (k == 1 ? nC : nF)) { kind = k; img = lo + sel; tk = got; break; }
            }
            __threadfence();                                         // acquire: the producers' data is visible from here on
            if (lane == 0) { s_tick[0] = kind; s_tick[1] = img; s_tick[2] = tk; }
        }
        __syncthreads();
        const int kind = s_tick[0], b = s_tick[1], g = s_tick[2];
        if (kind == 3) break;
        float2* img_ws = p.ws + (size_t)(b % R) * p.N * L;
        if (kind == 0) {
            // ------------------------------ forward rows ------------------------------
            const int y0 = g * rpt + w;
            if (prefetch && lane == 0 && b + 1 < p.planes) {
                // pull the same rows of the next image from HBM into L2 (whoever claims that ticket finds them there)
                for (int j = 0; j < rpw; ++j) k32_prefetch_row(p, b + 1, y0 + 8 * j);
            }
#pragma unroll 1
            for (int j = 0; j < rpw; ++j) {
                const int y = y0 + 8 * j;
                k32_row_fwd(p, line, tw, lane, b, y, img_ws + (size_t)y * L);
            }
            __syncthreads();
            if (t == 0) { __threadfence(); atomicAdd(done1 + b, rpt); }
        } else if (kind == 1) {
            // ------------------------------ column slabs ------------------------------
            k32_cols_items<true>(p, buf, buf, kz_s, tw, fold, 0, b * nslab + g * cq, b * nslab + (g + 1) * cq, 1, R);
            __syncthreads();
            if (t == 0) { __threadfence(); atomicAdd(done2 + b, cq); }
        } else {
            // ------------------------------ inverse rows + output stage ------------------------------
            const int y0 = g * rpt + w;
#pragma unroll 1
            for (int j = 0; j < rpw; ++j) {
                const int y = y0 + 8 * j;
                k32_row_inv(p, line, tw, lane, b, y, img_ws + (size_t)y * L);
            }
            __syncthreads();
            if (t == 0) { __threadfence(); atomicAdd(done3 + b, rpt); }
        }
    }
}

}  // namespace asmb

namespace asmb {

// ---------------------------------------------------------------------------------------------------
// Bulk-copy row kernels (default for FFT size 1024).  Measured on B200: the LDG/STG row kernels are bound by the
// number of global requests an SM can keep in flight (a second CTA per SM adds < 5 %, loads alone run at 4.2 TB/s),
// not by bandwidth or by the FFT.  Here every global access is ONE asynchronous bulk copy per row issued by one lane
// (cp.async.bulk, the TMA engine): HBM/L2 -> the warp's dense landing line (mbarrier completion), and the warp's
// staging line -> global (bulk_group).  The next row is requested as soon as the current one is in registers, so
// it lands during the transform; warps only execute shared-memory and FP instructions.
//   per warp: landing line (8 KB) | exchange line (8.25 KB, doubles as the dense store staging line)
//   12 warps per CTA, one CTA per SM.
// ---------------------------------------------------------------------------------------------------
constexpr int K32_BULK_WARPS = 12;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, unsigned bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// forward rows: in_mode COMPLEX or AMP_PHASE (any padding); other input modes use k32_rows_fwd
__global__ void __launch_bounds__(32 * K32_BULK_WARPS, 1) k32_rows_fwd_bulk(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LINE_B = K32_L * 8, XCH_B = K32_LP * 8;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    unsigned char* land = smem_raw + (size_t)w * (LINE_B + XCH_B);
    float2* xch = reinterpret_cast<float2*>(land + LINE_B);
    float2* tw = reinterpret_cast<float2*>(smem_raw + (size_t)K32_BULK_WARPS * (LINE_B + XCH_B));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + K32_TW);
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    if (t < K32_BULK_WARPS) mbar_init(bars + t, 1);
    fence_mbar_init();
    __syncthreads();
    uint64_t* bar = bars + w;
    const bool ap = p.in_mode == ASM_B200_IN_AMP_PHASE;
    const uint64_t pol_in = policy_evict_first(), pol_ws = policy_evict_normal();
    const int stride = gridDim.x * K32_BULK_WARPS;
    const unsigned row_bytes = (unsigned)p.N * 8u;                   // COMPLEX: N float2;  AMP_PHASE: N + N floats
    auto request = [&](int gline) {                                  // lane 0 only
        const size_t row = ((size_t)(plane0 + gline / p.N) * p.N + gline % p.N) * p.N;
        mbar_expect_tx(bar, row_bytes);
        if (ap) {
            bulk_load(land, (const float*)p.in0 + row, row_bytes / 2, bar, pol_in);
            bulk_load(land + row_bytes / 2, (const float*)p.in1 + row, row_bytes / 2, bar, pol_in);
        } else {
            bulk_load(land, (const float2*)p.in0 + row, row_bytes, bar, pol_in);
        }
    };
    int gline = blockIdx.x * K32_BULK_WARPS + w;
    if (lane == 0 && gline < nlines) request(gline);
    unsigned phase = 0;
    for (; gline < nlines; gline += stride) {
        mbar_wait(bar, phase);
        phase ^= 1u;
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            int x = lane + 32 * i - p.P;
            bool in = true;
            if (p.P != 0) { in = !p.adj || (x >= 0 && x < p.N); x = min(max(x, 0), p.N - 1); }
            if (ap) {
                const float a = reinterpret_cast<const float*>(land)[x];
                const float ph = reinterpret_cast<const float*>(land)[p.N + x] * p.in_scale;
                float sn, cs;
                sincos_full(ph, &sn, &cs);
                v[i] = in ? make_float2(a * cs, a * sn) : make_float2(0.f, 0.f);
            } else {
                v[i] = in ? reinterpret_cast<const float2*>(land)[x] : make_float2(0.f, 0.f);
            }
        }
        __syncwarp();                                                // the landing line is consumed
        if (lane == 0) {
            if (gline + stride < nlines) request(gline + stride);    // ... lands while this row is transformed
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has left the staging line
        }
        if (!(p.dbg & 256)) {
        fwd32_first(v);
        __syncwarp();
        sts16<RowLayout32, 5>(v, xch + lane);
        __syncwarp();
        lds16<RowLayout32, 0>(v, xch + 33 * lane);
        fwd32_table(v, tw + lane);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) xch[lane + 32 * i] = v[i];     // dense, natural frequency order
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const int wrow = (p.dbg & 2048) ? gline % (4 * p.N) : gline;   // timing experiment: keep the writes inside 32 MB
            bulk_store(p.ws + (size_t)wrow * K32_L, xch, LINE_B, pol_ws);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// inverse rows: out_mode INTENSITY (without the saved field) or COMPLEX; other output modes use k32_rows_inv
__global__ void __launch_bounds__(32 * K32_BULK_WARPS, 1) k32_rows_inv_bulk(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LINE_B = K32_L * 8, XCH_B = K32_LP * 8;
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    unsigned char* land = smem_raw + (size_t)w * (LINE_B + XCH_B);
    float2* xch = reinterpret_cast<float2*>(land + LINE_B);
    float2* tw = reinterpret_cast<float2*>(smem_raw + (size_t)K32_BULK_WARPS * (LINE_B + XCH_B));
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + K32_TW);
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    if (t < K32_BULK_WARPS) mbar_init(bars + t, 1);
    fence_mbar_init();
    __syncthreads();
    uint64_t* bar = bars + w;
    const uint64_t pol_out = policy_evict_first(), pol_ws = policy_evict_normal();
    const int stride = gridDim.x * K32_BULK_WARPS;
    const bool folding = p.adj && p.P > 0;
    const bool intensity = p.out_mode == ASM_B200_OUT_INTENSITY;
    auto request = [&](int gl) {
        mbar_expect_tx(bar, LINE_B);
        bulk_load(land, p.ws + (size_t)gl * K32_L, LINE_B, bar, pol_ws);
    };
    int gline = blockIdx.x * K32_BULK_WARPS + w;
    if (lane == 0 && gline < nlines) request(gline);
    unsigned phase = 0;
    for (; gline < nlines; gline += stride) {
        mbar_wait(bar, phase);
        phase ^= 1u;
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = reinterpret_cast<const float2*>(land)[lane + 32 * i];   // frequency lane + 32 i
        __syncwarp();
        if (!(p.dbg & 16)) {   // the intermediate row is dead: drop its dirty L2 lines instead of writing them back to HBM
            const char* src_row = reinterpret_cast<const char*>(p.ws + (size_t)gline * K32_L);
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(src_row + (size_t)lane * 128) : "memory");
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(src_row + (size_t)(lane + 32) * 128) : "memory");
        }
        if (lane == 0) {
            if (gline + stride < nlines) request(gline + stride);
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        inv32_first(v);
        __syncwarp();
        sts16<RowLayout32, 0>(v, xch + 33 * lane);
        __syncwarp();
        lds16<RowLayout32, 5>(v, xch + lane);
        inv32_table(v, tw + lane);                                   // v[i] = natural position lane + 32 i
        __syncwarp();
        const int img = gline / p.N, y = gline % p.N, plane = plane0 + img;
        float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
        if (folding) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int pos = lane + 32 * i;
                if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                if (pos >= p.P + p.N) { fr.x += v[i].x; fr.y += v[i].y; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fl.x += __shfl_xor_sync(0xffffffffu, fl.x, o); fl.y += __shfl_xor_sync(0xffffffffu, fl.y, o);
                fr.x += __shfl_xor_sync(0xffffffffu, fr.x, o); fr.y += __shfl_xor_sync(0xffffffffu, fr.y, o);
            }
        }
        // stage the cropped output row densely, then one bulk store
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int x = lane + 32 * i - p.P;
            if (x >= 0 && x < p.N) {
                float2 u = v[i];
                if (x == 0) { u.x += fl.x; u.y += fl.y; }
                if (x == p.N - 1) { u.x += fr.x; u.y += fr.y; }
                if (intensity) reinterpret_cast<float*>(xch)[x] = fmaf(u.x, u.x, u.y * u.y);
                else xch[x] = u;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const size_t row = ((size_t)plane * p.N + y) * p.N;
            if (intensity) bulk_store((float*)p.out0 + row, xch, (unsigned)p.N * 4u, pol_out);
            else bulk_store((float2*)p.out0 + row, xch, (unsigned)p.N * 8u, pol_out);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace asmb

// k32_cols.cuh -- column pass: FFT . H(z) . IFFT on 8-column slabs (plain and cp.async-pipelined)
// Part of the FFT-size-1024 path; included by k32.cuh (which is included by asm_b200.cu).
#pragma once

namespace asmb {

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

}  // namespace asmb

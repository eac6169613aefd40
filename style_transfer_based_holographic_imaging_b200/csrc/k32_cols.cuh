// k32_cols.cuh -- column pass of the FFT-size-1024 path: FFT . H(z) . IFFT on slabs of CC columns.
// Included by k32.cuh (which is included by asm_b200.cu).
//
// Persistent CTAs of 32 * CC threads (thread = column c = t % CC, 32 rows tl + 32 i with tl = t / CC), 16 / CC... CTAs
// per SM so that every SM holds 16 warps: CC = 8 -> 2 CTAs of 8 warps, CC = 4 -> 4 CTAs of 4 warps (four independent
// barrier domains per SM: the FFMA2-heavy, shared-memory-heavy and MUFU-heavy phases of different CTAs overlap).
// Per item (one slab of one sample):
//   the slab was copied global -> shared with cp.async into the dense first N rows of the exchange slab while the
//   previous item finished; dense rows -> registers, radix-32, exchange (padded [row][CC]), radix-32 with table
//   twiddles, x H(z) (kappa slab staged with cp.async during the previous inverse transform), inverse radix-32,
//   exchange, inverse radix-32, registers -> workspace (st.global.cg).  Crop / fold of the padding rows fused.
//   smem: slab [1056][CC] float2 | kappa [513][CC] double | half twiddle table [16][32] float2 | fold [2][CC] float2
#pragma once

namespace asmb {

template <int CC>
struct K32Cols {
    static constexpr int THREADS = 32 * CC;
    static constexpr int CTAS_PER_SM = 16 / CC;
    static constexpr int SLAB_ROWS = ColLayout32<CC>::rows(K32_L);        // 1056
    static constexpr size_t SMEM = (size_t)SLAB_ROWS * CC * 8 + (size_t)(K32_L / 2 + 1) * CC * 8 + (size_t)K32_TW * 8 + 2 * CC * 8;
};

template <int CC>
__device__ __forceinline__ void k32_stage_raw(float2* raw, const float2* img_ws, int col0, int nrows) {
    constexpr int Q = CC / 2;                                        // 16-byte pieces per row segment
    for (int j = threadIdx.x; j < nrows * Q; j += 32 * CC) {
        const int r = j / Q, q = j % Q;
        cp_async16(raw + r * CC + 2 * q, img_ws + (size_t)r * K32_L + col0 + 2 * q);
    }
}
template <int CC>
__device__ __forceinline__ void k32_stage_kz(double* kz_s, const double* kzt, int col0) {
    constexpr int Q = CC / 2;
    for (int j = threadIdx.x; j < (K32_L / 2 + 1) * Q; j += 32 * CC) {
        const int ru = j / Q, q = j % Q;
        cp_async16(kz_s + ru * CC + 2 * q, kzt + (size_t)ru * K32_L + col0 + 2 * q);
    }
}

template <int CC, bool PADDED>
__global__ void __launch_bounds__(32 * CC, 16 / CC) k32_cols(const Params p, int plane0, int nimg) {
    constexpr int L = K32_L, nslab = L / CC;
    using LAY = ColLayout32<CC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* slab = reinterpret_cast<float2*>(smem_raw);              // [1056][CC]; its first N rows double as the landing zone
    double* kz_s = reinterpret_cast<double*>(slab + K32Cols<CC>::SLAB_ROWS * CC);
    float2* tw = reinterpret_cast<float2*>(kz_s + (L / 2 + 1) * CC);
    float2* fold = tw + K32_TW;
    const int t = threadIdx.x, c = t % CC, tl = t / CC;
    const int N = PADDED ? p.N : L;
    const int total = nimg * nslab, step = gridDim.x;
    for (int i = t; i < K32_TW; i += 32 * CC) tw[i] = __ldg(p.tw + i);
    int wi = blockIdx.x;
    if (wi < total) {                                                // prologue: first slab + its kappa
        k32_stage_raw<CC>(slab, p.ws + (size_t)(wi / nslab) * N * L, (wi % nslab) * CC, N);
        k32_stage_kz<CC>(kz_s, p.kzt, (wi % nslab) * CC);
        cp_async_commit();
    }
    float2* col = slab + c;
    for (; wi < total; wi += step) {
        const int img = wi / nslab, slab_i = wi % nslab, plane = plane0 + img;
        const int col0 = slab_i * CC;
        const int nxt = wi + step;
        float2* img_ws = p.ws + (size_t)img * N * L;
        if (PADDED && t < 2 * CC) fold[t] = make_float2(0.f, 0.f);
        cp_async_wait<0>();                                          // this item's slab and kappa have landed
        __syncthreads();
        // ---- registers <- dense rows tl + 32 i (padding rows by clamp / zero) ----
        float2 v[32];
        if constexpr (!PADDED) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = slab[(tl + 32 * i) * CC + c];
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                int r = tl + 32 * i - p.P;
                if (p.adj) v[i] = (r >= 0 && r < N) ? slab[r * CC + c] : make_float2(0.f, 0.f);
                else { r = min(max(r, 0), N - 1); v[i] = slab[r * CC + c]; }
            }
        }
        __syncthreads();                                             // the dense rows are consumed
        const double cph = phase_constant_of(p, plane / p.C);

        // ---- forward column FFT ----
        fwd32_first(v);
        sts16<LAY, 5>(v, col + tl * CC);
        __syncthreads();
        lds16<LAY, 0>(v, col + 33 * tl * CC);
        fwd32_table(v, tw + tl);
        // ---- transfer function ----
        if (p.h_mode == H_DERIV) k32_apply_h<true, CC>(v, p, kz_s, c, tl, cph);
        else k32_apply_h<false, CC>(v, p, kz_s, c, tl, cph);
        // ---- inverse column FFT ----
        inv32_first(v);
        sts16<LAY, 0>(v, col + 33 * tl * CC);
        __syncthreads();                                             // every thread is done with kz_s too
        if (nxt < total && (nxt % nslab) != slab_i) k32_stage_kz<CC>(kz_s, p.kzt, (nxt % nslab) * CC);
        lds16<LAY, 5>(v, col + tl * CC);
        __syncthreads();                                             // the slab is dead from here on: land the next one in it
        if (nxt < total) k32_stage_raw<CC>(slab, p.ws + (size_t)(nxt / nslab) * N * L, (nxt % nslab) * CC, N);
        cp_async_commit();
        inv32_table(v, tw + tl);
        // ---- store rows [P, P+N) (crop); adjoint: fold the padding rows onto rows P and P+N-1 first ----
        float2* dst = img_ws + col0 + c;
        if constexpr (!PADDED) {
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(dst + (size_t)(tl + 32 * i) * L, v[i]);
        } else {
            float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
            if (p.adj) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pos = tl + 32 * i;
                    if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= p.P + N) { fr.x += v[i].x; fr.y += v[i].y; }
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
                if (r >= 0 && r < N) {
                    float2 u = v[i];
                    if (r == 0) { u.x += fl.x; u.y += fl.y; }
                    if (r == N - 1) { u.x += fr.x; u.y += fr.y; }
                    __stcg(dst + (size_t)r * L, u);
                }
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace asmb

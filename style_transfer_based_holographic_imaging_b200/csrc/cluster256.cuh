// cluster256.cuh -- FFT size 256 (256^2 unpadded: BASELINE config 2; 128^2 padded: the MNIST demo / test_data shape) with
// the WHOLE sample resident in the shared memory of a 4-CTA thread-block cluster: the two transposes of the 2-D transform
// go through distributed shared memory (DSMEM), the intermediate never touches L2.  One launch per call.
//
//   CTA `rank` of a cluster owns the source rows y = rank (mod 4): its row slab [N/4][272] holds their spectra.
//   phase 1  rows     HBM -> registers (input construction / padding fused) -> row FFT -> own row slab
//            cluster barrier
//   phase 2  columns  the CTA takes the 64 column positions [64 rank, 64 rank + 64) in two slabs of 32: every warp gathers
//            256-byte row segments from the four row slabs (ld.shared::cluster; 3 of 4 are remote) into a local column slab
//            [256][32], column FFT . H(z) . inverse FFT there (same code as the chunked column kernel), and the result
//            goes back to the owners' row slabs (st.shared::cluster)
//            cluster barrier
//   phase 3  rows     own row slab -> inverse row FFT -> crop / fold + output stage -> HBM
// The clusters are persistent (one per 4 SMs) and walk over the samples of the call.  HBM / L2 see the input once, the
// output once and the kappa entries of the CTA's 64 columns (66 KB per sample, L2 hits).  Every input / output / gradient
// mode of the chunked kernels is supported (same load16 / emit16 / fold code).
// Algorithmic traffic per sample: 12 N^2 B (forward) -- what the HBM roofline counts -- instead of 12 N^2 + 4 x 8 N M of
// L2 traffic for the three-pass pipeline.
// Included by asm_b200.cu after the generic kernels.
#pragma once

namespace asmb {

constexpr int C256_THREADS = 512, C256_CC = 32, C256_CLUSTER = 4;
constexpr int C256_LP = RowLayout::line_elems(256);                  // 272
constexpr size_t C256_SMEM = (size_t)64 * C256_LP * 8 + (size_t)256 * C256_CC * 8 + (size_t)make_layout(8).total * 8 + 2 * 64 * 8 + 16;

__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(const void* local, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
    return r;
}
__device__ __forceinline__ float2 dsmem_ld(uint32_t a) {
    float2 v;
    asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void dsmem_st(uint32_t a, float2 v) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}

__global__ void __cluster_dims__(C256_CLUSTER, 1, 1) __launch_bounds__(C256_THREADS, 1) k_cluster256(const Params p) {
    constexpr int n = 8, L = 256, TPL = 16, LP = C256_LP, CC = C256_CC, LPC = C256_THREADS / TPL;   // 32 lines per iteration
    constexpr TwLayout lay = make_layout(n);
    using CL = ColLayout<CC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* rowslab = reinterpret_cast<float2*>(smem_raw);           // [64][LP]: spectra of the rows this CTA owns
    float2* slab = rowslab + 64 * LP;                                // [L][CC]: column slab
    float2* tw = slab + L * CC;                                      // every table of n = 8
    float2* fold = tw + lay.total;                                   // rows: [LPC][2];  columns: [2][CC]
    const int t = threadIdx.x;
    const unsigned rank = cluster_rank();
    const int nclusters = gridDim.x / C256_CLUSTER, cid = blockIdx.x / C256_CLUSTER;
    const int N = p.N, P = p.P, rows_local = N / 4;
    for (int i = t; i < lay.total; i += C256_THREADS) tw[i] = __ldg(p.tw + i);
    __syncthreads();
    auto sync = [] { __syncthreads(); };
    const bool folding = p.adj && P > 0;

    for (int plane = cid; plane < p.planes; plane += nclusters) {
        // ---------------- phase 1: forward rows ----------------
        for (int it = 0; it * LPC < rows_local; ++it) {
            const int ll = t / TPL, tl = t % TPL, rl = it * LPC + ll, y = 4 * rl + (int)rank;
            float2 v[16];
            switch (p.in_mode) {
                case ASM_B200_IN_COMPLEX: load16<ASM_B200_IN_COMPLEX, TPL>(v, p, plane, y, tl); break;
                case ASM_B200_IN_AMP_PHASE: load16<ASM_B200_IN_AMP_PHASE, TPL>(v, p, plane, y, tl); break;
                case ASM_B200_IN_CONST_AMP_PHASE: load16<ASM_B200_IN_CONST_AMP_PHASE, TPL>(v, p, plane, y, tl); break;
                case ASM_B200_IN_SQRT_REAL: load16<ASM_B200_IN_SQRT_REAL, TPL>(v, p, plane, y, tl); break;
                case ASM_B200_IN_COT_FIELD: load16<ASM_B200_IN_COT_FIELD, TPL>(v, p, plane, y, tl); break;
                default: load16<ASM_B200_IN_REAL, TPL>(v, p, plane, y, tl); break;
            }
            float2* line = rowslab + rl * LP;
            fwd_line<n, RowLayout>(v, line, tw, tl, sync);
            sts16<RowLayout, 0>(v, line + RowLayout::base(thread_part(tl, 0)));
        }
        cluster_sync_all();                                          // every row spectrum of the sample is in some CTA's slab

        // ---------------- phase 2: columns ----------------
        const double cph = phase_constant_of(p, plane / p.C);
        for (int q = 0; q < 64 / CC; ++q) {
            const int c0 = 64 * (int)rank + CC * q;
            {   // gather: warp w takes rows w, w + 16, ...; its lanes read 32 consecutive column positions (256 bytes)
                const int c = t % CC, w = t / CC;
                const int src = RowLayout::phys(c0 + c);
                for (int y = w; y < N; y += C256_THREADS / CC)
                    slab[(size_t)(P + y) * CC + c] = dsmem_ld(dsmem_addr(rowslab + (y >> 2) * LP + src, (unsigned)(y & 3)));
                if (t < 2 * CC) fold[t] = make_float2(0.f, 0.f);
            }
            __syncthreads();
            const int c = t % CC, tl = t / CC;
            if (P > 0) {   // padding rows: replicate for the forward operator, zero for the adjoint
                const float2 top = slab[(size_t)P * CC + c], bot = slab[(size_t)(P + N - 1) * CC + c];
                const float2 zero = make_float2(0.f, 0.f);
                for (int r = tl; r < P; r += TPL) {
                    slab[(size_t)r * CC + c] = p.adj ? zero : top;
                    slab[(size_t)(P + N + r) * CC + c] = p.adj ? zero : bot;
                }
                __syncthreads();
            }
            float2* col = slab + c;
            float2 v[16];
            lds16<CL, n - 4>(v, col + CL::base(thread_part(tl, n - 4)));
            fwd_line<n, CL>(v, col, tw, tl, sync);
            {   // transfer function: register i holds column frequency u = Q + 16 i; kappa from the global table (L2)
                const int Q = fwd_q_from_hi(n, 0, tl);
                const double MAGIC = 6755399441055744.0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int u = Q + TPL * i;
                    const int ru = u <= L / 2 ? u : L - u;
                    const double kap = __ldg(p.kzt + (size_t)ru * L + c0 + c);
                    const double tt = kap * cph;
                    const double rr = tt - __dadd_rn(__dadd_rn(tt, MAGIC), -MAGIC);
                    float sn, cn;
                    __sincosf((float)rr * 6.283185307179586f, &sn, &cn);
                    if (p.h_mode == H_DERIV) v[i] = cmul_scaled(v[i], -sn, cn, (float)(kap * (6.283185307179586 * p.lambda) - p.kshift) * p.inv_m2);
                    else v[i] = cmul_scaled(v[i], cn, sn, p.inv_m2);
                }
            }
            inv_line<n, CL>(v, col, tw, tl, sync);
            sts16<CL, n - 4>(v, col + CL::base(thread_part(tl, n - 4)));
            if (folding) {   // adjoint of replicate padding along y: rows [0, P) onto row P, [P + N, M) onto row P + N - 1
                float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int pos = thread_part(tl, n - 4) | (i << (n - 4));
                    if (pos < P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= P + N) { fr.x += v[i].x; fr.y += v[i].y; }
                }
                atomicAdd(&fold[c].x, fl.x); atomicAdd(&fold[c].y, fl.y);
                atomicAdd(&fold[CC + c].x, fr.x); atomicAdd(&fold[CC + c].y, fr.y);
                __syncthreads();
                if (tl == 0) {
                    float2& r0 = slab[(size_t)P * CC + c];
                    float2& r1 = slab[(size_t)(P + N - 1) * CC + c];
                    r0.x += fold[c].x; r0.y += fold[c].y;
                    r1.x += fold[CC + c].x; r1.y += fold[CC + c].y;
                }
            }
            __syncthreads();
            {   // scatter rows [P, P + N) back to their owners
                const int cc = t % CC, w = t / CC;
                const int dst = RowLayout::phys(c0 + cc);
                for (int y = w; y < N; y += C256_THREADS / CC)
                    dsmem_st(dsmem_addr(rowslab + (y >> 2) * LP + dst, (unsigned)(y & 3)), slab[(size_t)(P + y) * CC + cc]);
            }
            __syncthreads();                                         // the column slab is reused by the next q
        }
        cluster_sync_all();                                          // every column result has reached its owner

        // ---------------- phase 3: inverse rows + output stage ----------------
        for (int it = 0; it * LPC < rows_local; ++it) {
            const int ll = t / TPL, tl = t % TPL, rl = it * LPC + ll, y = 4 * rl + (int)rank;
            float2* line = rowslab + rl * LP;
            if (folding && t < 2 * LPC) fold[t] = make_float2(0.f, 0.f);
            __syncthreads();
            float2 v[16];
            lds16<RowLayout, 0>(v, line + RowLayout::base(thread_part(tl, 0)));
            inv_line<n, RowLayout>(v, line, tw, tl, sync);           // positions tl + 16 i, natural order
            float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
            if (folding) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int pos = tl + TPL * i;
                    if (pos < P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= P + N) { fr.x += v[i].x; fr.y += v[i].y; }
                }
                atomicAdd(&fold[2 * ll].x, fl.x); atomicAdd(&fold[2 * ll].y, fl.y);
                atomicAdd(&fold[2 * ll + 1].x, fr.x); atomicAdd(&fold[2 * ll + 1].y, fr.y);
                __syncthreads();
                fl = fold[2 * ll]; fr = fold[2 * ll + 1];
            }
            float dot = 0.f;
            switch (p.out_mode) {
                case ASM_B200_OUT_COMPLEX: emit16<ASM_B200_OUT_COMPLEX, TPL>(v, p, plane, y, tl, fl, fr); break;
                case ASM_B200_OUT_INTENSITY: emit16<ASM_B200_OUT_INTENSITY, TPL>(v, p, plane, y, tl, fl, fr); break;
                case ASM_B200_OUT_ABS_ANGLE: emit16<ASM_B200_OUT_ABS_ANGLE, TPL>(v, p, plane, y, tl, fl, fr); break;
                case ASM_B200_OUT_REIM_CAT: emit16<ASM_B200_OUT_REIM_CAT, TPL>(v, p, plane, y, tl, fl, fr); break;
                case ASM_B200_OUT_ABSANG_CAT: emit16<ASM_B200_OUT_ABSANG_CAT, TPL>(v, p, plane, y, tl, fl, fr); break;
                case ASM_B200_OUT_GRAD_AP: emit16<ASM_B200_OUT_GRAD_AP, TPL>(v, p, plane, y, tl, fl, fr); break;
                default: dot = emit16<OUT_DOT, TPL>(v, p, plane, y, tl, fl, fr); break;
            }
            if (p.out_mode == OUT_DOT) {                             // the whole CTA works on one sample
                float s = dot;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const double K = p.z_f64 ? 6.283185307179586 : (double)6.2831854820251465f;
                if ((t & 31) == 0) atomicAdd((double*)p.out0 + plane / p.C, (double)s * K * p.inv_lambda);
            }
            __syncthreads();                                         // the lines are rewritten by the next sample's phase 1
        }
    }
}

}  // namespace asmb

// resident.cuh -- FFT sizes <= 512: the whole call as ONE persistent launch.
//
// The per-chunk pipeline (three launches per chunk) is launch bound for small transforms (256^2: ~800 launches of a few
// microseconds per fwd+adjoint step).  Here a GROUP of G co-resident CTAs (G = 1, 2, 4, 8) owns one sample at a time and
// takes it through all three passes itself; the sample's intermediate lives in the group's private slot of the
// workspace (N x M complex64, <= 2 MB), which is rewritten for every sample and therefore never leaves L2: the
// slots of all groups together are sized to ~74 MB of the 126 MB L2.  HBM sees the input once and the output once.
//   pass 1: rows  (HBM -> registers, input construction / padding fused) -> row FFT -> slot
//   pass 2: slabs of 16 columns (slot -> shared memory, 128-byte segments) -> column FFT . H(z) . IFFT -> slot
//   pass 3: rows  (slot -> shared memory) -> row IFFT -> crop / fold + output stage -> HBM
// Between the passes the CTAs of a group meet at a counter barrier in global memory (release / acquire at gpu scope);
// with G = 1 that is a plain __syncthreads().  Every group works on a different sample, and the 4 CTAs resident on an
// SM belong to different groups in different passes, so memory-bound and FP-bound phases overlap on every SM without
// any host-side scheduling.  All input / output modes of the chunked kernels are supported (same load16 / emit16).
// Included by asm_b200.cu after the generic kernels.
#pragma once

namespace asmb {

constexpr int RES_THREADS = 256;
constexpr int RES_CTAS_PER_SM = 4;

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all G CTAs of a group have passed barrier number `target / G` once the counter reaches `target`
__device__ __forceinline__ void group_barrier(int* counter, int target, int G) {
    __syncthreads();
    if (G > 1) {
        if (threadIdx.x == 0) {
            __threadfence();                              // this CTA's slot writes are visible device-wide ...
            atomicAdd(counter, 1);                        // ... before its arrival is
            while (ld_acquire_gpu(counter) < target) __nanosleep(40);
            __threadfence();
        }
        __syncthreads();
    }
}

template <int n>
struct ResidentCfg {
    static constexpr int L = 1 << n, TPL = L / 16, LPC = RES_THREADS / TPL, LP = RowLayout::line_elems(L);
    static constexpr int CC = n <= 8 ? 16 : 8;                            // columns per slab: CC * L / 16 threads transform one slab
    static constexpr int SLAB_T = CC * TPL, SPB = RES_THREADS / SLAB_T, NSLAB = L / CC, KZROWS = L / 2 + 1;
    static_assert(SLAB_T <= RES_THREADS, "a slab needs at most one CTA");
    static constexpr int NSI = (NSLAB + SPB - 1) / SPB;                   // slab iterations per sample
    static constexpr int BUF = LPC * LP > SPB * L * CC ? LPC * LP : SPB * L * CC;   // float2 elements (lines or slabs)
    static constexpr int NFOLD = 2 * LPC > SPB * 2 * CC ? 2 * LPC : SPB * 2 * CC;
    static constexpr size_t SMEM = (size_t)BUF * 8 + (size_t)SPB * KZROWS * CC * 8 + (size_t)make_layout(n).total * 8 + (size_t)NFOLD * 8;
};

template <int n>
__global__ void __launch_bounds__(RES_THREADS, RES_CTAS_PER_SM) k_resident(const Params p, int* ctl, int G) {
    using CFG = ResidentCfg<n>;
    constexpr int L = CFG::L, TPL = CFG::TPL, LPC = CFG::LPC, LP = CFG::LP;
    constexpr int CC = CFG::CC, SLAB_T = CFG::SLAB_T, SPB = CFG::SPB, NSLAB = CFG::NSLAB, KZROWS = CFG::KZROWS;
    constexpr TwLayout lay = make_layout(n);
    using CLAY = ColLayout<CC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);                    // row lines [LPC][LP]  or  column slabs [SPB][L][CC]
    double* kz_s = reinterpret_cast<double*>(buf + CFG::BUF);             // [SPB][KZROWS][CC]
    float2* tw = reinterpret_cast<float2*>(kz_s + SPB * KZROWS * CC);     // all tables of this n
    float2* fold = tw + lay.total;
    const int t = threadIdx.x;
    for (int i = t; i < lay.total; i += RES_THREADS) tw[i] = __ldg(p.tw + i);
    auto sync = [] { __syncthreads(); };

    const int gid = blockIdx.x / G, member = blockIdx.x % G, ngroups = gridDim.x / G;
    int* counter = ctl + gid;
    int epoch = 0;
    float2* slot = p.ws + (size_t)gid * p.N * L;                          // this group's L2-resident intermediate
    const int N = p.N, P = p.P;
    const int ntile = (N + LPC - 1) / LPC;
    const bool folding = p.adj && P > 0;

    // row-pass thread mapping
    const int ll = t / TPL, rtl = t % TPL;
    float2* line = buf + ll * LP;
    // column-pass thread mapping
    const int sg = t / SLAB_T, ts = t % SLAB_T, c = ts % CC, ctl_ = ts / CC;
    float2* slab = buf + (size_t)sg * L * CC;
    double* kz = kz_s + (size_t)sg * KZROWS * CC;
    float2* col = slab + c;
    __syncthreads();

    for (int plane = gid; plane < p.planes; plane += ngroups) {
        // ------------------------------------------------------------------------------------------------
        // pass 1: rows -> row FFT -> slot
        // ------------------------------------------------------------------------------------------------
        for (int tile = member; tile < ntile; tile += G) {
            const int y = tile * LPC + ll;
            const bool active = y < N;
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = make_float2(0.f, 0.f);
            if (active) {
                switch (p.in_mode) {
                    case ASM_B200_IN_COMPLEX: load16<ASM_B200_IN_COMPLEX, TPL>(v, p, plane, y, rtl); break;
                    case ASM_B200_IN_AMP_PHASE: load16<ASM_B200_IN_AMP_PHASE, TPL>(v, p, plane, y, rtl); break;
                    case ASM_B200_IN_CONST_AMP_PHASE: load16<ASM_B200_IN_CONST_AMP_PHASE, TPL>(v, p, plane, y, rtl); break;
                    case ASM_B200_IN_SQRT_REAL: load16<ASM_B200_IN_SQRT_REAL, TPL>(v, p, plane, y, rtl); break;
                    case ASM_B200_IN_COT_FIELD: load16<ASM_B200_IN_COT_FIELD, TPL>(v, p, plane, y, rtl); break;
                    default: load16<ASM_B200_IN_REAL, TPL>(v, p, plane, y, rtl); break;
                }
            }
            fwd_line<n, RowLayout>(v, line, tw, rtl, sync);
            sts16<RowLayout, 0>(v, line + RowLayout::base(thread_part(rtl, 0)));
            __syncthreads();
            if (active) {
                float2* dst = slot + (size_t)y * L;
#pragma unroll
                for (int k = 0; k < 16; ++k) __stcg(dst + rtl + TPL * k, line[RowLayout::phys(rtl) + RowLayout::off(k, n - 4)]);
            }
            __syncthreads();
        }
        group_barrier(counter, G * ++epoch, G);

        // ------------------------------------------------------------------------------------------------
        // pass 2: column slabs -> FFT . H(z) . IFFT -> slot
        // ------------------------------------------------------------------------------------------------
        const double cph = phase_constant_of(p, plane / p.C);
        for (int si = member; si < CFG::NSI; si += G) {
            const int slab_i = si * SPB + sg;
            const bool sactive = slab_i < NSLAB;
            if (sactive) {
                // slab rows are (8 CC)-byte segments of the slot (CC / 2 x 16 B); padding rows by clamp (forward) / zero (adjoint)
                constexpr int QP = CC / 2;
                for (int j = ts; j < L * QP; j += SLAB_T) {
                    const int r = j / QP, q = j % QP;
                    int sr = r - P;
                    const bool inside = sr >= 0 && sr < N;
                    sr = min(max(sr, 0), N - 1);
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (inside || !p.adj) val = __ldcg(reinterpret_cast<const float4*>(slot + (size_t)sr * L + slab_i * CC) + q);
                    reinterpret_cast<float4*>(slab + r * CC)[q] = val;
                }
                for (int j = ts; j < KZROWS * QP; j += SLAB_T) {
                    const int ru = j / QP, q = j % QP;
                    reinterpret_cast<float4*>(kz + ru * CC)[q] = __ldg(reinterpret_cast<const float4*>(p.kzt + (size_t)ru * L + slab_i * CC) + q);
                }
            }
            if (t < SPB * 2 * CC) fold[t] = make_float2(0.f, 0.f);
            __syncthreads();
            float2 v[16];
            lds16<CLAY, n - 4>(v, col + CLAY::base(thread_part(ctl_, n - 4)));
            fwd_line<n, CLAY>(v, col, tw, ctl_, sync);
            {   // transfer function: register i holds column frequency u = Q + (L/16) i
                const int Q = fwd_q_from_hi(n, 0, ctl_);
                const double MAGIC = 6755399441055744.0;                     // 1.5 * 2^52: round to nearest integer
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int u = Q + TPL * i;
                    const int ru = u <= L / 2 ? u : L - u;
                    const double kap = kz[ru * CC + c];
                    const double tt = kap * cph;
                    const double rr = tt - __dadd_rn(__dadd_rn(tt, MAGIC), -MAGIC);
                    float sn, cn;
                    __sincosf((float)rr * 6.283185307179586f, &sn, &cn);
                    if (p.h_mode == H_DERIV) v[i] = cmul_scaled(v[i], -sn, cn, (float)(kap * (6.283185307179586 * p.lambda) - p.kshift) * p.inv_m2);
                    else v[i] = cmul_scaled(v[i], cn, sn, p.inv_m2);
                }
            }
            inv_line<n, CLAY>(v, col, tw, ctl_, sync);
            sts16<CLAY, n - 4>(v, col + CLAY::base(thread_part(ctl_, n - 4)));
            if (folding) {
                // adjoint of replicate padding along rows: fold rows [0,P) onto row P and [P+N, M) onto row P+N-1
                float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int pos = thread_part(ctl_, n - 4) | (i << (n - 4));
                    if (pos < P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= P + N) { fr.x += v[i].x; fr.y += v[i].y; }
                }
                float2* f = fold + sg * 2 * CC;
                atomicAdd(&f[c].x, fl.x); atomicAdd(&f[c].y, fl.y);
                atomicAdd(&f[CC + c].x, fr.x); atomicAdd(&f[CC + c].y, fr.y);
                __syncthreads();
                if (ctl_ == 0) {
                    float2& r0 = slab[(size_t)P * CC + c];
                    float2& r1 = slab[(size_t)(P + N - 1) * CC + c];
                    r0.x += f[c].x; r0.y += f[c].y;
                    r1.x += f[CC + c].x; r1.y += f[CC + c].y;
                }
            }
            __syncthreads();
            if (sactive) {
                constexpr int QP = CC / 2;
                for (int j = ts; j < N * QP; j += SLAB_T) {
                    const int r = j / QP, q = j % QP;
                    __stcg(reinterpret_cast<float4*>(slot + (size_t)r * L + slab_i * CC) + q,
                           reinterpret_cast<const float4*>(slab + (size_t)(P + r) * CC)[q]);
                }
            }
            __syncthreads();
        }
        group_barrier(counter, G * ++epoch, G);

        // ------------------------------------------------------------------------------------------------
        // pass 3: slot -> row IFFT -> crop / fold + output stage
        // ------------------------------------------------------------------------------------------------
        for (int tile = member; tile < ntile; tile += G) {
            const int y = tile * LPC + ll;
            const bool active = y < N;
            if (active) {
                const float2* src = slot + (size_t)y * L;
#pragma unroll
                for (int k = 0; k < 16; ++k) line[RowLayout::phys(rtl) + RowLayout::off(k, n - 4)] = __ldcg(src + rtl + TPL * k);
            }
            if (folding && t < 2 * LPC) fold[t] = make_float2(0.f, 0.f);
            __syncthreads();
            float2 v[16];
            lds16<RowLayout, 0>(v, line + RowLayout::base(thread_part(rtl, 0)));
            inv_line<n, RowLayout>(v, line, tw, rtl, sync);
            float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
            if (folding) {
                if (active) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int pos = rtl + TPL * i;
                        if (pos < P) { fl.x += v[i].x; fl.y += v[i].y; }
                        if (pos >= P + N) { fr.x += v[i].x; fr.y += v[i].y; }
                    }
                    atomicAdd(&fold[2 * ll].x, fl.x); atomicAdd(&fold[2 * ll].y, fl.y);
                    atomicAdd(&fold[2 * ll + 1].x, fr.x); atomicAdd(&fold[2 * ll + 1].y, fr.y);
                }
                __syncthreads();
                fl = fold[2 * ll]; fr = fold[2 * ll + 1];
            }
            float dot = 0.f;
            if (active) {
                switch (p.out_mode) {
                    case ASM_B200_OUT_COMPLEX: emit16<ASM_B200_OUT_COMPLEX, TPL>(v, p, plane, y, rtl, fl, fr); break;
                    case ASM_B200_OUT_INTENSITY: emit16<ASM_B200_OUT_INTENSITY, TPL>(v, p, plane, y, rtl, fl, fr); break;
                    case ASM_B200_OUT_ABS_ANGLE: emit16<ASM_B200_OUT_ABS_ANGLE, TPL>(v, p, plane, y, rtl, fl, fr); break;
                    case ASM_B200_OUT_REIM_CAT: emit16<ASM_B200_OUT_REIM_CAT, TPL>(v, p, plane, y, rtl, fl, fr); break;
                    case ASM_B200_OUT_ABSANG_CAT: emit16<ASM_B200_OUT_ABSANG_CAT, TPL>(v, p, plane, y, rtl, fl, fr); break;
                    case ASM_B200_OUT_GRAD_AP: emit16<ASM_B200_OUT_GRAD_AP, TPL>(v, p, plane, y, rtl, fl, fr); break;
                    default: dot = emit16<OUT_DOT, TPL>(v, p, plane, y, rtl, fl, fr); break;
                }
            }
            if (p.out_mode == OUT_DOT) {   // a tile belongs to one sample: warp reduce, one double atomic per warp
                float s = active ? dot : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const double K = p.z_f64 ? 6.283185307179586 : (double)6.2831854820251465f;
                if ((t & 31) == 0) atomicAdd((double*)p.out0 + plane / p.C, (double)s * K * p.inv_lambda);
            }
            __syncthreads();   // the line buffers (and fold) are reused by the next tile
        }
        group_barrier(counter, G * ++epoch, G);   // the slot is rewritten by the next sample's pass 1
    }
}

}  // namespace asmb

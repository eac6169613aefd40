// k32_persistent.cuh -- opt-in alternatives: cp.async row ring, static-role persistent kernel (k32_mega), uniform-worker persistent kernel (k32_flow)
// Part of the FFT-size-1024 path; included by k32.cuh (which is included by asm_b200.cu).
#pragma once

namespace asmb {

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

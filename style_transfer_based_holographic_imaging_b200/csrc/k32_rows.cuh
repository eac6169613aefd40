// k32_rows.cuh -- row passes of the FFT-size-1024 path: TMA bulk-copy kernels (default) and register-landing
// LDG/STG kernels (every other input / output mode, unaligned buffers).
// Included by k32.cuh (which is included by asm_b200.cu).
#pragma once

namespace asmb {

#ifdef ASM_B200_TUNING   /* the round-1 row-major kernels: A/B builds only (the product runs k32t.cuh) */
// ---------------------------------------------------------------------------------------------------
// Register-landing row kernels: one warp per row, 8 warps per CTA, 2 CTAs per SM.  All input / output modes.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void k32_row_fwd(const Params& p, float2* line, const float2* tw, int lane, int plane, int y,
                                            float2* dst_row) {
    float2 v[32];
    switch (p.in_mode) {
        case ASM_B200_IN_COMPLEX: load32<ASM_B200_IN_COMPLEX>(v, p, plane, y, lane); break;
        case ASM_B200_IN_AMP_PHASE: load32<ASM_B200_IN_AMP_PHASE>(v, p, plane, y, lane); break;
        case ASM_B200_IN_CONST_AMP_PHASE: load32<ASM_B200_IN_CONST_AMP_PHASE>(v, p, plane, y, lane); break;
        case ASM_B200_IN_SQRT_REAL: load32<ASM_B200_IN_SQRT_REAL>(v, p, plane, y, lane); break;
        case ASM_B200_IN_COT_FIELD: load32<ASM_B200_IN_COT_FIELD>(v, p, plane, y, lane); break;
        default: load32<ASM_B200_IN_REAL>(v, p, plane, y, lane); break;
    }
    fwd32_first(v);                                              // digit of bits 5..9 (positions lane + 32 i)
    sts16<RowLayout32, 5>(v, line + lane);
    __syncwarp();
    lds16<RowLayout32, 0>(v, line + 33 * lane);                  // positions 32 lane + i
    fwd32_table(v, tw + lane);
    __syncwarp();                                                // line is rewritten by the next row
    float2* dst = dst_row + lane;                                // frequency lane + 32 i
#pragma unroll
    for (int i = 0; i < 32; ++i) __stcg(dst + 32 * i, v[i]);
}

__global__ void __launch_bounds__(32 * K32_ROW_WARPS, K32_ROW_CTAS) k32_rows_fwd(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* lines = reinterpret_cast<float2*>(smem_raw);             // [K32_ROW_WARPS][K32_LP]
    float2* tw = lines + K32_ROW_WARPS * K32_LP;                     // [16][32]
    const int t = threadIdx.x, w = t >> 5, lane = t & 31;
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw + i);
    __syncthreads();
    float2* line = lines + w * K32_LP;
    const bool prefetch = k32_prefetch_ok(p);
    for (int gline = blockIdx.x * K32_ROW_WARPS + w; gline < nlines; gline += gridDim.x * K32_ROW_WARPS) {
        const int img = gline / p.N, y = gline % p.N;
        // pull this warp's NEXT source row from HBM into L2 while this one is transformed
        const int nxt = gline + gridDim.x * K32_ROW_WARPS;
        if (prefetch && lane == 0 && nxt < nlines) k32_prefetch_row(p, plane0 + nxt / p.N, nxt % p.N);
        k32_row_fwd(p, line, tw, lane, plane0 + img, y, p.ws + (size_t)gline * K32_L);
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
    {
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

#endif  // ASM_B200_TUNING

// ---------------------------------------------------------------------------------------------------
// Bulk-copy row kernels (default for FFT size 1024).  Every global access is ONE asynchronous bulk copy per row issued
// by one lane (cp.async.bulk, the TMA engine): HBM/L2 -> the warp's dense landing line (mbarrier completion), and the
// warp's staging line -> global (bulk_group).  The next row is requested as soon as the current one is in registers,
// so it lands during the transform; warps only execute shared-memory and FP instructions.
//   per warp: landing line (8 KB) | exchange line (8.25 KB, doubles as the dense store staging line)
//   K32_BULK_WARPS warps per CTA, one CTA per SM.
// Templates strip the padding / mode logic from the hot (unpadded) instantiations.
// ---------------------------------------------------------------------------------------------------
#ifndef K32_BULK_WARPS_DEF
#define K32_BULK_WARPS_DEF 6
#endif
constexpr int K32_BULK_WARPS = K32_BULK_WARPS_DEF;

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

// The landing line is rewritten by the NEXT bulk load (an async-proxy write), which is issued right after the current row
// has been read out of it (generic-proxy reads).  __syncwarp() orders the lanes' instruction streams, but a shared-memory
// load that is still queued in the MIO pipeline has not read its data yet -- nothing in the code uses the values before the
// request is issued -- and with a co-resident CTA hammering shared memory such a load can still be pending when the next
// row lands (observed with two CTAs per SM: about one wrong row in 1e4).  A cross-proxy fence executed by every lane
// before the barrier orders the reads (it waits for the thread's outstanding shared-memory accesses) before the request.
__device__ __forceinline__ void loads_landed(const float2 (&v)[32]) {
    (void)v;
    fence_proxy_async();
}

// IN: 0 complex64 rows, 1 amplitude + phase rows, 2 constant amplitude + phase rows
// REGST: the transformed row leaves straight from registers (coalesced st.global.cg) instead of staging + bulk store
template <int IN, bool PADDED, bool REGST = false>
__global__ void __launch_bounds__(32 * K32_BULK_WARPS, 12 / K32_BULK_WARPS > 0 ? 12 / K32_BULK_WARPS : 1) k32_rows_fwd_bulk(const Params p, int plane0, int nlines) {
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
    const uint64_t pol_in = policy_evict_first(), pol_ws = policy_evict_normal();
    const int stride = gridDim.x * K32_BULK_WARPS;
    const int N = PADDED ? p.N : K32_L;
    const unsigned row_bytes = (unsigned)N * (IN == 2 ? 4u : 8u);    // complex: N float2;  amp + phase: N + N floats;  const amp: N floats
    const float amp0 = IN == 2 ? __ldg((const float*)p.in0) : 0.f;
    auto request = [&](int gline) {                                  // lane 0 only
        const size_t row = ((size_t)(plane0 + gline / N) * N + gline % N) * N;
        mbar_expect_tx(bar, row_bytes);
        if constexpr (IN == 1) {
            bulk_load(land, (const float*)p.in0 + row, row_bytes / 2, bar, pol_in);
            bulk_load(land + row_bytes / 2, (const float*)p.in1 + row, row_bytes / 2, bar, pol_in);
        } else if constexpr (IN == 2) {
            bulk_load(land, (const float*)p.in1 + row, row_bytes, bar, pol_in);
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
            int x = lane + 32 * i;
            bool in = true;
            if constexpr (PADDED) { x -= p.P; in = !p.adj || (x >= 0 && x < N); x = min(max(x, 0), N - 1); }
            float2 val;
            if constexpr (IN == 0) {
                val = reinterpret_cast<const float2*>(land)[x];
            } else {
                const float a = IN == 1 ? reinterpret_cast<const float*>(land)[x] : amp0;
                const float ph = reinterpret_cast<const float*>(land)[(IN == 1 ? N : 0) + x] * p.in_scale;
                float sn, cs;
                sincos_reduced(ph, &sn, &cs);
                val = make_float2(a * cs, a * sn);
            }
            v[i] = (!PADDED || in) ? val : make_float2(0.f, 0.f);
        }
        loads_landed(v);
        __syncwarp();                                                // the landing line is consumed
        if (lane == 0) {
            if (gline + stride < nlines) request(gline + stride);    // ... lands while this row is transformed
            if constexpr (!REGST) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has left the staging line
        }
        fwd32_first(v);
        __syncwarp();
        sts16<RowLayout32, 5>(v, xch + lane);
        __syncwarp();
        lds16<RowLayout32, 0>(v, xch + 33 * lane);
        fwd32_table(v, tw + lane);
        __syncwarp();
        if constexpr (REGST) {
            float2* dst = p.ws + (size_t)gline * K32_L + lane;      // frequency lane + 32 i
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(dst + 32 * i, v[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) xch[lane + 32 * i] = v[i]; // dense, natural frequency order
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                bulk_store(p.ws + (size_t)gline * K32_L, xch, LINE_B, pol_ws);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// inverse rows: out_mode INTENSITY (without the saved field) or COMPLEX; other output modes use k32_rows_inv
template <bool INTENSITY, bool PADDED, bool REGST = false>
__global__ void __launch_bounds__(32 * K32_BULK_WARPS, 12 / K32_BULK_WARPS > 0 ? 12 / K32_BULK_WARPS : 1) k32_rows_inv_bulk(const Params p, int plane0, int nlines) {
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
    const int N = PADDED ? p.N : K32_L;
    const bool folding = PADDED && p.adj;
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
        loads_landed(v);
        __syncwarp();
        {   // the intermediate row is dead: drop its dirty L2 lines instead of writing them back to HBM
            const char* src_row = reinterpret_cast<const char*>(p.ws + (size_t)gline * K32_L);
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(src_row + (size_t)lane * 128) : "memory");
            asm volatile("discard.global.L2 [%0], 128;" ::"l"(src_row + (size_t)(lane + 32) * 128) : "memory");
        }
        if (lane == 0) {
            if (gline + stride < nlines) request(gline + stride);
            if constexpr (!REGST) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        inv32_first(v);
        __syncwarp();
        sts16<RowLayout32, 0>(v, xch + 33 * lane);
        __syncwarp();
        lds16<RowLayout32, 5>(v, xch + lane);
        inv32_table(v, tw + lane);                                   // v[i] = natural position lane + 32 i
        __syncwarp();
        const int img = gline / N, y = gline % N, plane = plane0 + img;
        if constexpr (REGST && !PADDED) {
            const size_t row = ((size_t)plane * N + y) * N + lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if constexpr (INTENSITY) __stcs((float*)p.out0 + row + 32 * i, fmaf(v[i].x, v[i].x, v[i].y * v[i].y));
                else __stcs((float2*)p.out0 + row + 32 * i, v[i]);
            }
            continue;
        }
        if constexpr (!PADDED) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if constexpr (INTENSITY) reinterpret_cast<float*>(xch)[lane + 32 * i] = fmaf(v[i].x, v[i].x, v[i].y * v[i].y);
                else xch[lane + 32 * i] = v[i];
            }
        } else {
            float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
            if (folding) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pos = lane + 32 * i;
                    if (pos < p.P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= p.P + N) { fr.x += v[i].x; fr.y += v[i].y; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    fl.x += __shfl_xor_sync(0xffffffffu, fl.x, o); fl.y += __shfl_xor_sync(0xffffffffu, fl.y, o);
                    fr.x += __shfl_xor_sync(0xffffffffu, fr.x, o); fr.y += __shfl_xor_sync(0xffffffffu, fr.y, o);
                }
            }
            // stage the cropped output row densely
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int x = lane + 32 * i - p.P;
                if (x >= 0 && x < N) {
                    float2 u = v[i];
                    if (x == 0) { u.x += fl.x; u.y += fl.y; }
                    if (x == N - 1) { u.x += fr.x; u.y += fr.y; }
                    if constexpr (INTENSITY) reinterpret_cast<float*>(xch)[x] = fmaf(u.x, u.x, u.y * u.y);
                    else xch[x] = u;
                }
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const size_t row = ((size_t)plane * N + y) * N;
            if constexpr (INTENSITY) bulk_store((float*)p.out0 + row, xch, (unsigned)N * 4u, pol_out);
            else bulk_store((float2*)p.out0 + row, xch, (unsigned)N * 8u, pol_out);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace asmb

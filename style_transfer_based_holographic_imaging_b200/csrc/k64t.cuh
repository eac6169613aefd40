// k64t.cuh -- the FFT-size-2048 path with a TRANSPOSED intermediate (the k32t.cuh recipe with the 2 x 1024 warp-pair
// transform of k64.cuh): every pass works on contiguous 16 KB lines that are private to one WARP PAIR, the two transposes
// are done by the TMA engine on 32-byte segments (tiles of 4 rows), and H(z) is evaluated in double-float fp32.
//
//   wsT[img][u'][y]   u' = row-pass frequency in SPLIT order (u' = 1024 h + m holds frequency 2 m + h), y = source row (N)
//
//   pass 1  k64t_rows_fwd : a CTA = 4 warp pairs = 4 consecutive source rows.  One 66 KB region serves as the pairs' landing
//                           lines (TMA bulk copy of the source row), their exchange lines, and finally as the tile
//                           [2048 u'][4 y] (SWIZZLE_32B) that 8 TMA tensor stores of 256 u' x 4 y send to the workspace.
//   pass 2  k64t_lines    : a warp pair per line u': the 16 KB line (and the kappa row of |u|) lands in the pair's exchange
//                           lines by TMA bulk copies issued during the previous line's last stage; forward 2048-point
//                           transform over y, x H(u, v), inverse transform, coalesced stores back in place.  Only pair
//                           barriers (bar.sync id, 64): 8 independent pairs per SM.
//   pass 3  k64t_rows_inv : mirror of pass 1 (8 TMA tensor loads land the tile, inverse transform, |U|^2 / complex64 rows
//                           leave by one bulk store per half); quads of adjacent groups drop their consumed workspace
//                           lines from L2 (discard.global.L2).
// Used when both row passes qualify for TMA bulk copies (complex64 / amplitude+phase / constant amplitude in, complex64 /
// |U|^2 out, 16-byte aligned rows); every other mode takes k64.cuh + the generic row kernels as before.
// Reference semantics: utils/Angular_Spectrum_Method.py:7-36.  Included by asm_b200.cu after k64.cuh.
#pragma once

namespace asmb {

constexpr int K64T_PAIRS = 4;                                        // warp pairs per CTA (all three passes)
constexpr int K64T_PAIR_B = 2 * K32_LP * 8;                          // 16 896: two exchange lines = one dense 16 KB landing line
constexpr int K64T_TILE_B = 2048 * 4 * 8;                            // [2048 u'][4 y] complex64
constexpr int KAP2_STRIDE = 1032;                                    // entries per row of the symmetric kappa table (1025 used)
constexpr unsigned KAP2_ROW_B = 1026 * 8;                            // bytes copied per line (multiple of 16)
constexpr size_t K64T_ROWS_SMEM = 1024 + (size_t)K64T_PAIRS * K64T_PAIR_B + (size_t)K32_TW * 8 + 64;
constexpr size_t K64T_LINES_SMEM = (size_t)K64T_PAIRS * (K64T_PAIR_B + KAP2_STRIDE * 8) + (size_t)K32_TW * 8 + 64;

// tables: half twiddle table of the 1024-point transform + kappa2[ru][rv] = (hi, lo) for 0 <= ru, rv <= 1024
__global__ void k64t_setup(float2* tw, float2* kap2, double s2, double inv_2pi_lambda) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    for (int e = gtid; e < K32_TW; e += gsz) {
        const int ent = e / 32, Q = e % 32;
        int m = 1;
        while (m < 5 && ent >= (1 << (m - 1))) ++m;
        const int u = ent - (m == 1 ? 0 : (1 << (m - 2)));
        const int D = 32 << m, x = Q + 32 * u;
        float sn, cs;
        sincospif(2.0f * (float)x / (float)D, &sn, &cs);
        tw[e] = make_float2(cs, -sn);
    }
    for (int idx = gtid; idx < 1025 * KAP2_STRIDE; idx += gsz) {
        const int ru = idx / KAP2_STRIDE, rv = idx % KAP2_STRIDE;
        const double kk = (double)ru * ru + (double)rv * rv;
        const double arg = fma(-s2, kk, 1.0);
        const double kap = (arg > 0.0 ? sqrt(arg) : 0.0) * inv_2pi_lambda;
        const float hi = (float)kap;
        kap2[idx] = make_float2(hi, (float)(kap - (double)hi));
    }
}

// element index (float2 units) of (u', yy) in the SWIZZLE_32B tile: 32-byte row u', 16-byte chunk (yy / 2) ^ ((u' / 4) & 1)
__device__ __forceinline__ int k64t_tile_idx(int u, int yy) { return u * 4 + ((((yy >> 1) ^ (u >> 2)) & 1) << 1) + (yy & 1); }

// H(u, v) on the registers of half h: register i holds frequency v = 2 (lane + 32 i) + h (see k32t_apply_h)
template <bool DERIV>
__device__ __forceinline__ void k64t_apply_h(float2 (&v)[32], const Params& p, const float2* kap_s, int lane, int h, float c_hi, float c_lo) {
    const float MAGIC = 12582912.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int f = 2 * (lane + 32 * i) + h;
        const float2 k = kap_s[f <= 1024 ? f : 2048 - f];
        const float pp = __fmul_rn(k.x, c_hi);
        float s = __fmaf_rn(k.x, c_hi, -pp);
        s = __fmaf_rn(k.y, c_hi, s);
        s = __fmaf_rn(k.x, c_lo, s);
        const float kk = __fadd_rn(__fadd_rn(pp, MAGIC), -MAGIC);
        const float r = __fadd_rn(__fadd_rn(pp, -kk), s);
        float sn, cn;
        __sincosf(r * 6.283185307179586f, &sn, &cn);
        if constexpr (DERIV) {
            const double kz_l = ((double)k.x + (double)k.y) * (6.283185307179586 * p.lambda);
            v[i] = cmul_scaled(v[i], -sn, cn, (float)(kz_l - p.kshift) * p.inv_m2);
        } else {
            v[i] = cmul_scaled(v[i], cn, sn, p.inv_m2);
        }
    }
}

// forward 2048-point transform of a warp pair: `fetch(pos)` returns element pos of the (padded) line; on return register i
// of half h holds frequency 2 (lane + 32 i) + h.  `xch` = this half's exchange line; `pair_bar` = the pair's barrier id.
template <class Fetch>
__device__ __forceinline__ void k64t_forward(float2 (&v)[32], Fetch fetch, float2* xch, const float2* tw, int lane, int h, int pair_bar) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float2 lo = fetch(lane + 32 * i), hi = fetch(lane + 32 * i + 1024);
        v[i] = h == 0 ? make_float2(lo.x + hi.x, lo.y + hi.y) : make_float2(lo.x - hi.x, lo.y - hi.y);
    }
    pair_barrier(pair_bar);                                          // both halves have consumed the landing line
    if (h == 1) k64_twiddle64<-1>(v);
    fwd32_first(v);
    sts16<RowLayout32, 5>(v, xch + lane);
    __syncwarp();
    lds16<RowLayout32, 0>(v, xch + 33 * lane);
    if (h == 1) k64_twiddle2048<-1>(v);
    fwd32_table(v, tw + lane);
}

// inverse 2048-point transform: register i of half h holds frequency 2 (lane + 32 i) + h on entry and position
// 1024 h + lane + 32 i on return.  Both exchange lines are free again after the second pair barrier.
__device__ __forceinline__ void k64t_inverse(float2 (&v)[32], float2* xch, const float2* xch_other, const float2* tw, int lane, int h, int pair_bar) {
    inv32_first(v);
    if (h == 1) k64_twiddle2048<1>(v);
    __syncwarp();
    sts16<RowLayout32, 0>(v, xch + 33 * lane);
    __syncwarp();
    lds16<RowLayout32, 5>(v, xch + lane);
    inv32_table(v, tw + lane);
    if (h == 1) k64_twiddle64<1>(v);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) xch[lane + 32 * i] = v[i];
    pair_barrier(pair_bar);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float2 o = xch_other[lane + 32 * i];
        v[i] = h == 0 ? make_float2(v[i].x + o.x, v[i].y + o.y) : make_float2(o.x - v[i].x, o.y - v[i].y);
    }
    pair_barrier(pair_bar);
}

// ---------------------------------------------------------------------------------------------------
// pass 1.  IN: 0 complex64 rows, 1 amplitude + phase rows, 2 constant amplitude + phase rows
// ---------------------------------------------------------------------------------------------------
template <int IN, bool PADDED>
__global__ void __launch_bounds__(64 * K64T_PAIRS, 2)
k64t_rows_fwd(const Params p, const __grid_constant__ CUtensorMap tmapT, int plane0, int ngroups) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float2* tile = reinterpret_cast<float2*>(base);                  // [2048 u'][4 y], aliases the lines
    const int t = threadIdx.x, pr = t >> 6, h = (t >> 5) & 1, lane = t & 31;
    unsigned char* land = base + (size_t)pr * K64T_PAIR_B;           // dense source row of the pair (<= 16 KB)
    float2* xch = reinterpret_cast<float2*>(land) + h * K32_LP;
    float2* tw = reinterpret_cast<float2*>(base + (size_t)K64T_PAIRS * K64T_PAIR_B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + K32_TW);
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw32 + i);
    if (t < K64T_PAIRS) mbar_init(bars + t, 1);
    fence_mbar_init();
    uint64_t* bar = bars + pr;
    const uint64_t pol_in = policy_evict_first();
    const int N = PADDED ? p.N : K64_L;
    const unsigned row_bytes = (unsigned)N * (IN == 2 ? 4u : 8u);
    const float amp0 = IN == 2 ? __ldg((const float*)p.in0) : 0.f;
    auto fetch = [&](int pos) -> float2 {
        int x = pos;
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
        return (!PADDED || in) ? val : make_float2(0.f, 0.f);
    };
    unsigned phase = 0;
    float2* tcol = tile + k64t_tile_idx(1024 * h + lane, pr);        // u' = 1024 h + lane + 32 i: (u' / 4) & 1 = (lane / 4) & 1
    for (int g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const int gline = g * 4 + pr;
        __syncthreads();                                             // the region is free (the previous tile has left it)
        if (h == 0 && lane == 0) {
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
            if (g + gridDim.x < ngroups) {                           // HBM -> L2 for the next group's row
                const int nxt = (g + gridDim.x) * 4 + pr;
                const size_t nrow = ((size_t)(plane0 + nxt / N) * N + nxt % N) * N;
                if constexpr (IN == 1) { l2_prefetch_bulk((const float*)p.in0 + nrow, row_bytes / 2); l2_prefetch_bulk((const float*)p.in1 + nrow, row_bytes / 2); }
                else if constexpr (IN == 2) l2_prefetch_bulk((const float*)p.in1 + nrow, row_bytes);
                else l2_prefetch_bulk((const float2*)p.in0 + nrow, row_bytes);
            }
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        float2 v[32];
        k64t_forward(v, fetch, xch, tw, lane, h, 1 + pr);            // v[i] = frequency 2 (lane + 32 i) + h
        __syncthreads();                                             // every pair is done with its lines: the region becomes the tile
#pragma unroll
        for (int i = 0; i < 32; ++i) tcol[i * 128] = v[i];
        fence_proxy_async();
        __syncthreads();
        if (t == 0) {
            const int img = (g * 4) / N, y0 = (g * 4) % N;
#pragma unroll
            for (int k = 0; k < 8; ++k) tma_store_3d(&tmapT, tile + k * 1024, 2 * y0, 256 * k, img);
            tma_commit();
            tma_wait_read0();
        }
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------
// pass 2: one warp pair per line u' of the transposed workspace
// ---------------------------------------------------------------------------------------------------
template <bool PADDED>
__global__ void __launch_bounds__(64 * K64T_PAIRS, 2) k64t_lines(const Params p, int plane0, int nlines) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int PB = K64T_PAIR_B + KAP2_STRIDE * 8;
    const int t = threadIdx.x, pr = t >> 6, h = (t >> 5) & 1, lane = t & 31;
    float2* land = reinterpret_cast<float2*>(smem_raw + (size_t)pr * PB);        // dense line of the pair = its two exchange lines
    float2* xch = land + h * K32_LP;
    const float2* xch_other = land + (1 - h) * K32_LP;
    float2* kap_s = reinterpret_cast<float2*>(smem_raw + (size_t)pr * PB + K64T_PAIR_B);
    float2* tw = reinterpret_cast<float2*>(smem_raw + (size_t)K64T_PAIRS * PB);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tw + K32_TW);
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw32 + i);
    if (t < K64T_PAIRS) mbar_init(bars + t, 1);
    fence_mbar_init();
    __syncthreads();
    uint64_t* bar = bars + pr;
    const uint64_t pol = policy_evict_normal();
    const int N = PADDED ? p.N : K64_L, P = PADDED ? p.P : 0;
    const float2* kap2 = reinterpret_cast<const float2*>(p.kzt);
    // lines sorted by |u| (u = 2 m' + h' for line u' = 1024 h' + m'): a pair's consecutive lines share the kappa row
    const int nimg = nlines >> 11;
    auto line_of = [&](int s) {
        int img, u;
        if (s < nimg) { img = s; u = 0; }
        else if (s >= nlines - nimg) { img = s - (nlines - nimg); u = 1024; }
        else {
            const int q = s - nimg, r = 1 + q / (2 * nimg), rem = q - (r - 1) * 2 * nimg;
            img = rem >> 1;
            u = (rem & 1) ? 2048 - r : r;
        }
        return (img << 11) | ((u & 1) << 10) | (u >> 1);             // u' = 1024 (u mod 2) + u / 2
    };
    auto ru_of = [](int line) { const int up = line & 2047, u = 2 * (up & 1023) + (up >> 10); return u <= 1024 ? u : 2048 - u; };
    auto request = [&](int line, bool with_kappa) {                  // lane 0 of half 0 only
        mbar_expect_tx(bar, (unsigned)N * 8u + (with_kappa ? KAP2_ROW_B : 0u));
        bulk_load(land, p.ws + (size_t)line * N, (unsigned)N * 8u, bar, pol);
        if (with_kappa) bulk_load(kap_s, kap2 + (size_t)ru_of(line) * KAP2_STRIDE, KAP2_ROW_B, bar, pol);
    };
    const int npairs = gridDim.x * K64T_PAIRS, pg = blockIdx.x * K64T_PAIRS + pr;
    int s = (int)((long long)pg * nlines / npairs);
    const int s_end = (int)((long long)(pg + 1) * nlines / npairs);
    if (h == 0 && lane == 0 && s < s_end) request(line_of(s), true);
    unsigned phase = 0;
    for (; s < s_end; ++s) {
        const int line = line_of(s);
        const int img = line >> 11;
        float2* row = p.ws + (size_t)line * N;
        float c_hi, c_lo;
        k32t_phase_constant(p, (plane0 + img) / p.C, &c_hi, &c_lo);
        mbar_wait(bar, phase);
        phase ^= 1u;
        auto fetch = [&](int pos) -> float2 {
            if constexpr (!PADDED) return land[pos];
            const int y = pos - P;
            if (p.adj) return (y >= 0 && y < N) ? land[y] : make_float2(0.f, 0.f);
            return land[min(max(y, 0), N - 1)];
        };
        float2 v[32];
        k64t_forward(v, fetch, xch, tw, lane, h, 1 + pr);
        if (p.h_mode == H_DERIV) k64t_apply_h<true>(v, p, kap_s, lane, h, c_hi, c_lo);
        else k64t_apply_h<false>(v, p, kap_s, lane, h, c_hi, c_lo);
        k64t_inverse(v, xch, xch_other, tw, lane, h, 1 + pr);        // v[i] = position 1024 h + lane + 32 i
        fence_proxy_async();                                         // line and kappa reads are done before the next line lands
        pair_barrier(1 + pr);
        if (h == 0 && lane == 0 && s + 1 < s_end) {
            const int nxt = line_of(s + 1);
            request(nxt, ru_of(nxt) != ru_of(line));
        }
        if constexpr (!PADDED) {
#pragma unroll
            for (int i = 0; i < 32; ++i) __stcg(row + 1024 * h + lane + 32 * i, v[i]);
        } else {
            float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
            if (p.adj) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int pos = 1024 * h + lane + 32 * i;
                    if (pos < P) { fl.x += v[i].x; fl.y += v[i].y; }
                    if (pos >= P + N) { fr.x += v[i].x; fr.y += v[i].y; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    fl.x += __shfl_xor_sync(0xffffffffu, fl.x, o); fl.y += __shfl_xor_sync(0xffffffffu, fl.y, o);
                    fr.x += __shfl_xor_sync(0xffffffffu, fr.x, o); fr.y += __shfl_xor_sync(0xffffffffu, fr.y, o);
                }
            }
            // P <= 1024 <= P + N: half 0 owns the fold onto y = 0, half 1 the fold onto y = N - 1
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int y = 1024 * h + lane + 32 * i - P;
                if (y >= 0 && y < N) {
                    float2 o = v[i];
                    if (y == 0) { o.x += fl.x; o.y += fl.y; }
                    if (y == N - 1) { o.x += fr.x; o.y += fr.y; }
                    __stcg(row + y, o);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// pass 3.  OUT: 0 complex64, 1 |U|^2 (one TMA bulk store per half)
// ---------------------------------------------------------------------------------------------------
template <int OUT, bool PADDED>
__global__ void __launch_bounds__(64 * K64T_PAIRS, 2)
k64t_rows_inv(const Params p, const __grid_constant__ CUtensorMap tmapT, int plane0, int ngroups) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float2* tile = reinterpret_cast<float2*>(base);
    const int t = threadIdx.x, pr = t >> 6, h = (t >> 5) & 1, lane = t & 31;
    float2* xch0 = reinterpret_cast<float2*>(base + (size_t)pr * K64T_PAIR_B);
    float2* xch = xch0 + h * K32_LP;
    const float2* xch_other = xch0 + (1 - h) * K32_LP;
    float2* tw = reinterpret_cast<float2*>(base + (size_t)K64T_PAIRS * K64T_PAIR_B);
    uint64_t* bar = reinterpret_cast<uint64_t*>(tw + K32_TW);
    for (int i = t; i < K32_TW; i += blockDim.x) tw[i] = __ldg(p.tw32 + i);
    if (t == 0) mbar_init(bar, 1);
    fence_mbar_init();
    const uint64_t pol_out = policy_evict_first();
    const int N = PADDED ? p.N : K64_L, P = PADDED ? p.P : 0;
    const bool folding = PADDED && p.adj;
    const int nquads = ngroups >> 2;                                 // N / 4 is a multiple of 4: quads = 16 consecutive rows
    const float2* tcol = tile + k64t_tile_idx(1024 * h + lane, pr);
    unsigned phase = 0;
    for (int it = 0; blockIdx.x + (it >> 2) * gridDim.x < nquads; ++it) {
        const int g = 4 * (blockIdx.x + (it >> 2) * gridDim.x) + (it & 3);
        __syncthreads();                                             // the region is free (every output row has left its line)
        if (t == 0) {
            const int img = (g * 4) / N, y0 = (g * 4) % N;
            mbar_expect_tx(bar, K64T_TILE_B);
#pragma unroll
            for (int k = 0; k < 8; ++k) tma_load_3d(tile + k * 1024, &tmapT, bar, 2 * y0, 256 * k, img);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        float2 v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = tcol[i * 128];           // frequency 2 (lane + 32 i) + h of row 4 g + pr
        __syncthreads();                                             // the tile is in registers: the region becomes the lines
        if ((it & 3) == 3) {   // the 128-byte workspace lines of these 16 rows are dead
            const int img = ((g - 3) * 4) / N, y0 = ((g - 3) * 4) % N;
            const char* a = reinterpret_cast<const char*>(p.ws + ((size_t)img * K64_L + t) * N + y0);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("discard.global.L2 [%0], 128;" ::"l"(a + (size_t)j * 256 * N * 8) : "memory");
        }
        k64t_inverse(v, xch, xch_other, tw, lane, h, 1 + pr);        // v[i] = position 1024 h + lane + 32 i
        const int gline = g * 4 + pr;
        const int img = gline / N, y = gline % N, plane = plane0 + img;
        float2 fl = make_float2(0.f, 0.f), fr = make_float2(0.f, 0.f);
        if (folding) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int pos = 1024 * h + lane + 32 * i;
                if (pos < P) { fl.x += v[i].x; fl.y += v[i].y; }
                if (pos >= P + N) { fr.x += v[i].x; fr.y += v[i].y; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                fl.x += __shfl_xor_sync(0xffffffffu, fl.x, o); fl.y += __shfl_xor_sync(0xffffffffu, fl.y, o);
                fr.x += __shfl_xor_sync(0xffffffffu, fr.x, o); fr.y += __shfl_xor_sync(0xffffffffu, fr.y, o);
            }
        }
        const int x0 = PADDED ? (h == 0 ? 0 : 1024 - P) : 1024 * h;
        const int cnt = PADDED ? (h == 0 ? 1024 - P : P + N - 1024) : 1024;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int x = 1024 * h + lane + 32 * i - P;
            if (!PADDED || (x >= 0 && x < N)) {
                float2 u = v[i];
                if (PADDED && x == 0) { u.x += fl.x; u.y += fl.y; }
                if (PADDED && x == N - 1) { u.x += fr.x; u.y += fr.y; }
                if constexpr (OUT == 1) reinterpret_cast<float*>(xch)[x - x0] = fmaf(u.x, u.x, u.y * u.y);
                else xch[x - x0] = u;
            }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            if (cnt > 0) {
                const size_t row = ((size_t)plane * N + y) * N + x0;
                if constexpr (OUT == 1) bulk_store((float*)p.out0 + row, xch, (unsigned)cnt * 4u, pol_out);
                else bulk_store((float2*)p.out0 + row, xch, (unsigned)cnt * 8u, pol_out);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            tma_wait_read0();
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace asmb

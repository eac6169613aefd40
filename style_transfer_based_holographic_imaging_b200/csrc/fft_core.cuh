// fft_core.cuh -- in-register / in-shared-memory FFT building blocks for the ASM propagator (sm_100a).
//
// A line of L = 2^n complex points (n = 5..12) is transformed by threads that own 16 points each.  The
// position bits of the line are split into fields: an optional top field A of a = n % 4 bits (radix 2^a)
// and nf = n / 4 radix-16 fields F_{nf-1} .. F_0 (F_0 = lowest 4 bits).  Every stage transforms ONE field
// in place (a thread reads 16 positions, does radix-2 DIT levels on them in registers, writes the same
// 16 positions back), so exactly one barrier separates two stages and no data is ever reordered:
//   forward  : fields top -> bottom;  X[u] ends up at the position whose fields hold the digits of u in
//              reversed order (digit-reversed output).
//   inverse  : fields bottom -> top, consuming that digit-reversed order and producing natural order.
// All twiddles are folded into the butterflies (a +- w*b = 6 FMA-class ops); the first stage of each
// direction has compile-time twiddles, later stages read W_{S*2^m}^{Q + S*u} from a table indexed by the
// thread's accumulated digit Q (S = product of the radices already done).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace asmb {

// ---------------------------------------------------------------------------------------------------
// plan / table layout (host + device)
// ---------------------------------------------------------------------------------------------------
struct TwLayout {
    int fwd[3];   // offset of the forward table of field F_f (-1: compile-time twiddles)
    int inv[3];   // offset of the inverse table of field F_f (-1: compile-time twiddles)
    int invA;     // offset of the inverse table of the top field A (-1: none)
    int fwd_end;  // forward tables occupy [0, fwd_end)
    int total;    // inverse tables occupy [fwd_end, total)
};

__host__ __device__ constexpr int ipow16(int e) { return e <= 0 ? 1 : 16 * ipow16(e - 1); }

__host__ __device__ constexpr TwLayout make_layout(int n) {
    TwLayout t{{-1, -1, -1}, {-1, -1, -1}, -1, 0, 0};
    const int a = n % 4, nf = n / 4, r = 1 << a;
    int off = 0;
    for (int f = nf - 1; f >= 0; --f) {
        const int S = r * ipow16(nf - 1 - f);
        if (S > 1) { t.fwd[f] = off; off += 15 * S; }
    }
    t.fwd_end = off;
    for (int f = 1; f < nf; ++f) { t.inv[f] = off; off += 15 * ipow16(f); }
    if (a > 0) { t.invA = off; off += (r - 1) * ((1 << n) / r); }
    t.total = off;
    return t;
}

// bank-conflict swizzle of a row-major line: XOR the low nibble of the position with the next nibble.
// An involution that permutes positions inside aligned groups of 16 only.
__host__ __device__ __forceinline__ constexpr int swz(int p) { return p ^ ((p >> 4) & 15); }

// ---------------------------------------------------------------------------------------------------
// butterflies
// ---------------------------------------------------------------------------------------------------
// (x, y) <- (x + w*y, x - w*y), w = (wr, wi).  6 FMA-class instructions.
__device__ __forceinline__ void bf(float2& x, float2& y, float wr, float wi) {
    const float o1r = fmaf(-y.y, wi, fmaf(y.x, wr, x.x));
    const float o1i = fmaf(y.y, wr, fmaf(y.x, wi, x.y));
    y.x = fmaf(2.f, x.x, -o1r);
    y.y = fmaf(2.f, x.y, -o1i);
    x.x = o1r;
    x.y = o1i;
}
__device__ __forceinline__ void bf_one(float2& x, float2& y) {  // w = 1
    const float2 t = y;
    y.x = x.x - t.x; y.y = x.y - t.y;
    x.x = x.x + t.x; x.y = x.y + t.y;
}
template <bool INV>
__device__ __forceinline__ void bf_quarter(float2& x, float2& y) {  // w = -i (forward) / +i (inverse)
    const float2 t = y;
    if (!INV) { y.x = x.x - t.y; y.y = x.y + t.x; x.x = x.x + t.y; x.y = x.y - t.x; }
    else      { y.x = x.x + t.y; y.y = x.y - t.x; x.x = x.x - t.y; x.y = x.y + t.x; }
}

// cos/sin(2*pi*k/16), k = 0..7
__device__ __forceinline__ constexpr float c16(int k) {
    return k == 0 ? 1.f : k == 1 ? 0.92387953251128674f : k == 2 ? 0.70710678118654752f : k == 3 ? 0.38268343236508977f
         : k == 4 ? 0.f : k == 5 ? -0.38268343236508977f : k == 6 ? -0.70710678118654752f : -0.92387953251128674f;
}
__device__ __forceinline__ constexpr float s16(int k) {
    return k == 0 ? 0.f : k == 1 ? 0.38268343236508977f : k == 2 ? 0.70710678118654752f : k == 3 ? 0.92387953251128674f
         : k == 4 ? 1.f : k == 5 ? 0.92387953251128674f : k == 6 ? 0.70710678118654752f : 0.38268343236508977f;
}

__host__ __device__ constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// ---------------------------------------------------------------------------------------------------
// one in-register stage: radix R = 2^E DIT over the register sub-field  i = (k << SH) | g,  k in [0,R)
// for every g in [0, 1<<SH)  (SH = 0 for a full radix-16 field; SH = 4-a for the top field A).
//   CONST : compile-time twiddles W_{2^m}^u (first stage of a direction)
//   else  : tw[e * tws + q(g)] with e = 2^{m-1}-1+u; q(g) = q0 + g * qg  (group-dependent digit for field A)
// ---------------------------------------------------------------------------------------------------
template <int E, int SH, bool INV, bool CONST>
__device__ __forceinline__ void stage16(float2 (&v)[16], const float2* __restrict__ tw, int tws, int q0, int qg) {
    constexpr int R = 1 << E;
    constexpr int G = 1 << SH;
#pragma unroll
    for (int g = 0; g < G; ++g) {
        float2 a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = v[(bitrev(j, E) << SH) | g];
        const float2* twg = CONST ? nullptr : tw + (q0 + g * qg);
#pragma unroll
        for (int m = 1; m <= E; ++m) {
            const int half = 1 << (m - 1);
#pragma unroll
            for (int b = 0; b < R; b += 2 * half) {
#pragma unroll
                for (int u = 0; u < half; ++u) {
                    float2& x = a[b + u];
                    float2& y = a[b + u + half];
                    if constexpr (CONST) {
                        const int k16 = u * (16 >> m);  // angle index in sixteenths of a turn, 0..7
                        if (k16 == 0) bf_one(x, y);
                        else if (k16 == 4) bf_quarter<INV>(x, y);
                        else bf(x, y, c16(k16), INV ? s16(k16) : -s16(k16));
                    } else {
                        const float2 w = twg[(half - 1 + u) * tws];
                        bf(x, y, w.x, w.y);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) v[(j << SH) | g] = a[j];
    }
}

// ---------------------------------------------------------------------------------------------------
// thread <-> position maps.  tl in [0, L/16) is the thread's index inside its line.
// ---------------------------------------------------------------------------------------------------
// window with low bit w0: position of register i
__device__ __forceinline__ int win_pos(int tl, int w0, int i) {
    const int lo = tl & ((1 << w0) - 1);
    const int hi = tl >> w0;
    return (hi << (w0 + 4)) | (i << w0) | lo;
}

// forward accumulated digit for field F_f: digits of the already transformed (higher) fields in reversed
// weight order: A (weight 1), F_{nf-1} (weight r), F_{nf-2} (weight 16 r), ...
template <int n>
__device__ __forceinline__ int fwd_q(int tl, int f) {
    constexpr int a = n % 4, nf = n / 4;
    const int hi = tl >> (4 * f);
    const int nb = n - 4 * f - 4;  // bits in hi
    int Q = 0, w = 1;
    if (a > 0) { Q = hi >> (nb - a); w = 1 << a; }
#pragma unroll
    for (int g = nf - 1; g > f; --g) {
        Q += ((hi >> (4 * (g - f - 1))) & 15) * w;
        w *= 16;
    }
    return Q;
}

// ---------------------------------------------------------------------------------------------------
// stage drivers on a shared-memory line.  ADDR(pos) maps a line position to a float2 index in `sm`.
// ---------------------------------------------------------------------------------------------------
template <class Addr>
__device__ __forceinline__ void lds16(float2 (&v)[16], const float2* sm, Addr addr, int tl, int w0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = sm[addr(win_pos(tl, w0, i))];
}
template <class Addr>
__device__ __forceinline__ void sts16(const float2 (&v)[16], float2* sm, Addr addr, int tl, int w0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) sm[addr(win_pos(tl, w0, i))] = v[i];
}

// first forward stage (window n-4): constant twiddles; radix 2^a if a > 0 else 16
template <int n>
__device__ __forceinline__ void fwd_first(float2 (&v)[16]) {
    constexpr int a = n % 4;
    if constexpr (a > 0) stage16<a, 4 - a, false, true>(v, nullptr, 0, 0, 0);
    else                 stage16<4, 0, false, true>(v, nullptr, 0, 0, 0);
}
// forward table stage of field F_f (window 4f); tw = base of all tables in shared memory
template <int n, int f>
__device__ __forceinline__ void fwd_field(float2 (&v)[16], const float2* tw, int tl) {
    constexpr TwLayout lay = make_layout(n);
    constexpr int a = n % 4, nf = n / 4;
    constexpr int S = (1 << a) * ipow16(nf - 1 - f);
    static_assert(lay.fwd[f] >= 0, "field has no forward table");
    stage16<4, 0, false, false>(v, tw + lay.fwd[f], S, fwd_q<n>(tl, f), 0);
}
// first inverse stage: field F_0, constant twiddles
__device__ __forceinline__ void inv_first(float2 (&v)[16]) { stage16<4, 0, true, true>(v, nullptr, 0, 0, 0); }
// inverse table stage of field F_f, f >= 1 (window 4f): S' = 16^f, Q' = low bits of the position
template <int n, int f>
__device__ __forceinline__ void inv_field(float2 (&v)[16], const float2* tw, int tl) {
    constexpr TwLayout lay = make_layout(n);
    constexpr int S = ipow16(f);
    static_assert(lay.inv[f] >= 0, "field has no inverse table");
    stage16<4, 0, true, false>(v, tw + lay.inv[f], S, tl & (S - 1), 0);
}
// inverse stage of the top field A (window n-4): radix 2^a, S' = L / 2^a, Q'(g) = tl + g * L/16
template <int n>
__device__ __forceinline__ void inv_top(float2 (&v)[16], const float2* tw, int tl) {
    constexpr TwLayout lay = make_layout(n);
    constexpr int a = n % 4;
    if constexpr (a > 0) {
        constexpr int S = (1 << n) >> a;
        stage16<a, 4 - a, true, false>(v, tw + lay.invA, S, tl, (1 << n) / 16);
    }
}

// frequency index held by register i of a thread after the LAST forward stage (field F_0, window 0):
// u = Q + (L/16) * i with Q = fwd_q(tl, 0).
// frequency index of physical (swizzled) row position c of a row that went through the forward row pass:
template <int n>
__host__ __device__ __forceinline__ int freq_of_pos(int pos) {
    constexpr int a = n % 4, nf = n / 4;
    int Q = 0, w = 1;
    if (a > 0) { Q = pos >> (n - a); w = 1 << a; }
    for (int g = nf - 1; g >= 0; --g) {
        Q += ((pos >> (4 * g)) & 15) * w;
        w *= 16;
    }
    return Q;
}

}  // namespace asmb

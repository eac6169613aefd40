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
// direction has compile-time twiddles, later stages read W_{S*2^m}^{Q + S*u} from a table (S = product of
// the radices already done, Q = the thread's accumulated digit).
//
// Addressing: a thread's 16 positions are  T | (i << w0)  (T = thread part, w0 = low bit of its 4-bit
// window).  Both shared-memory layouts used here are additive in disjoint bit fields, so the address is
// base(T) + a COMPILE-TIME offset(i): every LDS/STS uses an immediate offset, no index arithmetic.
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

// forward accumulated digit Q of a thread whose already-transformed (higher) position bits are `hi`
// (field F_f is next; f = -1: all fields done, hi = the whole position).  Digits in reversed weight
// order: A (weight 1), F_{nf-1} (weight 2^a), F_{nf-2} (weight 16 * 2^a), ...
__host__ __device__ constexpr int fwd_q_from_hi(int n, int f, int hi) {
    const int a = n % 4, nf = n / 4;
    const int nb = n - 4 * f - 4;
    int Q = 0, w = 1;
    if (a > 0) { Q = hi >> (nb - a); w = 1 << a; }
    for (int g = nf - 1; g > f; --g) {
        Q += ((hi >> (4 * (g - f - 1))) & 15) * w;
        w *= 16;
    }
    return Q;
}
// frequency index stored at position `pos` of a line after the complete forward pass
__host__ __device__ constexpr int freq_of_pos(int n, int pos) { return fwd_q_from_hi(n, -1, pos); }

// ---------------------------------------------------------------------------------------------------
// shared-memory layouts
// ---------------------------------------------------------------------------------------------------
// Row layout: one pad element after every 16 (bank-conflict free for every stage: a thread's 16 contiguous
// positions are 17 elements from its neighbour's).  phys() is additive over disjoint bit fields.
struct RowLayout {
    __host__ __device__ static constexpr int phys(int pos) { return pos + (pos >> 4); }
    __host__ __device__ static constexpr int line_elems(int L) { return L + (L >> 4); }
    __host__ __device__ static constexpr int off(int i, int w0) { return phys(i << w0); }
    __device__ static __forceinline__ int base(int T) { return phys(T); }
};
// Column layout: [position][CC columns], dense (what a TMA box of CC columns x rows lands as)
template <int CC>
struct ColLayout {
    __host__ __device__ static constexpr int off(int i, int w0) { return (i << w0) * CC; }
    __device__ static __forceinline__ int base(int T) { return T * CC; }
};

// thread part of the position for the 4-bit window starting at bit w0
__device__ __forceinline__ int thread_part(int tl, int w0) {
    const int lo = tl & ((1 << w0) - 1);
    const int hi = tl >> w0;
    return (hi << (w0 + 4)) | lo;
}

template <class LAY, int W0, int PT>
__device__ __forceinline__ void lds16(float2 (&v)[PT], const float2* base) {
#pragma unroll
    for (int i = 0; i < PT; ++i) v[i] = base[LAY::off(i, W0)];
}
template <class LAY, int W0, int PT>
__device__ __forceinline__ void sts16(const float2 (&v)[PT], float2* base) {
#pragma unroll
    for (int i = 0; i < PT; ++i) base[LAY::off(i, W0)] = v[i];
}

// ---------------------------------------------------------------------------------------------------
// butterflies
// ---------------------------------------------------------------------------------------------------
// Packed fp32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2): a complex number (re, im) is ONE 64-bit register pair.
// ptxas folds the half swaps / sign patterns written below as repacks into operand modifiers of the packed
// instruction (e.g. `FFMA2 R2, -R26.F32x2.LO_HI.NP, R3.F32, R28.F32x2.HI_LO`: swapped halves, (-,+) signs, scalar
// broadcast), so a complex butterfly with twiddle is 3 instructions and a trivial one 2, with no extra moves.
typedef unsigned long long u64x;
__device__ __forceinline__ u64x pk2(float lo, float hi) { u64x r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64x pk2(float2 v) { return pk2(v.x, v.y); }
__device__ __forceinline__ float2 upk2(u64x v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64x fma2(u64x a, u64x b, u64x c) { u64x d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64x add2(u64x a, u64x b) { u64x d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64x mul2(u64x a, u64x b) { u64x d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// v <- v * (hr + i hi) * scale  (packed: 1 FMUL2 for the scaling, FMUL2 + FFMA2 for the complex product)
__device__ __forceinline__ float2 cmul_scaled(float2 v, float hr, float hi, float scale) {
    const float2 h = upk2(mul2(pk2(hr, hi), pk2(scale, scale)));
    return upk2(fma2(pk2(v), pk2(h.x, h.x), mul2(pk2(-v.y, v.x), pk2(h.y, h.y))));
}

#ifndef ASM_B200_SCALAR_BF
// (x, y) <- (x + w*y, x - w*y), w = (wr, wi).  3 packed FMA instructions.
__device__ __forceinline__ void bf(float2& x, float2& y, float wr, float wi) {
    const u64x X = pk2(x);
    const u64x t = fma2(pk2(y), pk2(wr, wr), X);                       // x + wr * (yr, yi)
    const u64x o1 = fma2(pk2(-y.y, y.x), pk2(wi, wi), t);              //   + wi * (-yi, yr)
    const float2 o1f = upk2(o1);
    y = upk2(fma2(pk2(2.f, 2.f), X, pk2(-o1f.x, -o1f.y)));             // 2x - o1
    x = o1f;
}
__device__ __forceinline__ void bf_one(float2& x, float2& y) {  // w = 1
    const u64x X = pk2(x);
    const float2 t = y;
    y = upk2(add2(X, pk2(-t.x, -t.y)));
    x = upk2(add2(X, pk2(t)));
}
template <bool INV>
__device__ __forceinline__ void bf_quarter(float2& x, float2& y) {  // w = -i (forward) / +i (inverse)
    const u64x X = pk2(x);
    const float2 t = y;
    if constexpr (!INV) { y = upk2(add2(X, pk2(-t.y, t.x))); x = upk2(add2(X, pk2(t.y, -t.x))); }
    else                { y = upk2(add2(X, pk2(t.y, -t.x))); x = upk2(add2(X, pk2(-t.y, t.x))); }
}
#else
// scalar reference versions (A/B builds only): 6 / 4 FMA-class instructions
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
    if constexpr (!INV) { y.x = x.x - t.y; y.y = x.y + t.x; x.x = x.x + t.y; x.y = x.y - t.x; }
    else                { y.x = x.x + t.y; y.y = x.y - t.x; x.x = x.x - t.y; x.y = x.y + t.x; }
}
#endif

// cos/sin(2*pi*k/32), k = 1..15 (k = 0, 8 are handled by the trivial butterflies)
template <int K> struct Rot32 { static constexpr float c = 1.f, s = 0.f; };
template <> struct Rot32<1> { static constexpr float c = 0.98078528040323043f, s = 0.19509032201612825f; };
template <> struct Rot32<2> { static constexpr float c = 0.92387953251128674f, s = 0.38268343236508978f; };
template <> struct Rot32<3> { static constexpr float c = 0.83146961230254524f, s = 0.55557023301960218f; };
template <> struct Rot32<4> { static constexpr float c = 0.70710678118654757f, s = 0.70710678118654746f; };
template <> struct Rot32<5> { static constexpr float c = 0.55557023301960229f, s = 0.83146961230254524f; };
template <> struct Rot32<6> { static constexpr float c = 0.38268343236508984f, s = 0.92387953251128674f; };
template <> struct Rot32<7> { static constexpr float c = 0.19509032201612833f, s = 0.98078528040323043f; };
template <> struct Rot32<9> { static constexpr float c = -0.19509032201612819f, s = 0.98078528040323043f; };
template <> struct Rot32<10> { static constexpr float c = -0.38268343236508973f, s = 0.92387953251128674f; };
template <> struct Rot32<11> { static constexpr float c = -0.55557023301960196f, s = 0.83146961230254546f; };
template <> struct Rot32<12> { static constexpr float c = -0.70710678118654746f, s = 0.70710678118654757f; };
template <> struct Rot32<13> { static constexpr float c = -0.83146961230254535f, s = 0.55557023301960218f; };
template <> struct Rot32<14> { static constexpr float c = -0.92387953251128674f, s = 0.38268343236508989f; };
template <> struct Rot32<15> { static constexpr float c = -0.98078528040323043f, s = 0.19509032201612861f; };

__host__ __device__ constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// one butterfly of DIT level M (block size 2^M) at block offset B, index U -- everything compile time
// CONST: compile-time twiddle W_{2^M}^U (conjugated when INV).  Otherwise the twiddle is read from the table
// as stored, or conjugated on the fly when CONJ (a table stored for the forward direction then serves both).
// HT (half table): level M stores only its first q = max(1, 2^{M-2}) twiddles (entries off(M) .. off(M)+q-1,
// off(1) = 0, off(M) = 2^{M-2}); the upper half is -i times the lower half (W^{S 2^{M-2}} = W_4 = -i), which costs
// nothing: x +- (-i w) y uses (w_i, -w_r) as the two scalar operands of the same three FFMA2s.
template <int R, int M, int B, int U, bool INV, bool CONST, int S, bool CONJ, bool HT>
__device__ __forceinline__ void bfly(float2 (&a)[R], const float2* __restrict__ tw) {
    constexpr int half = 1 << (M - 1);
    float2& x = a[B + U];
    float2& y = a[B + U + half];
    if constexpr (CONST) {
        constexpr int k32 = U * (32 >> M);  // angle in 32nds of a turn, 0..15
        if constexpr (k32 == 0) bf_one(x, y);
        else if constexpr (k32 == 8) bf_quarter<INV>(x, y);
        else bf(x, y, Rot32<k32>::c, INV ? Rot32<k32>::s : -Rot32<k32>::s);
    } else if constexpr (HT) {
        constexpr int q = half >= 2 ? half / 2 : 1;
        constexpr int off = M == 1 ? 0 : (1 << (M - 2));
        if constexpr (U < q) {
            const float2 w = tw[(off + U) * S];
            bf(x, y, w.x, CONJ ? -w.y : w.y);
        } else {
            const float2 w = tw[(off + U - q) * S];          // -i w0 = (w0.y, -w0.x); conjugated: (w0.y, +w0.x)
            bf(x, y, w.y, CONJ ? w.x : -w.x);
        }
    } else {
        const float2 w = tw[(half - 1 + U) * S];
        bf(x, y, w.x, CONJ ? -w.y : w.y);
    }
}
template <int R, int M, int I, bool INV, bool CONST, int S, bool CONJ, bool HT>
__device__ __forceinline__ void level_iter(float2 (&a)[R], const float2* __restrict__ tw) {
    // I enumerates the R/2 butterflies of level M: block = I / half, u = I % half
    if constexpr (I < R / 2) {
        constexpr int half = 1 << (M - 1);
        bfly<R, M, (I / half) * 2 * half, I % half, INV, CONST, S, CONJ, HT>(a, tw);
        level_iter<R, M, I + 1, INV, CONST, S, CONJ, HT>(a, tw);
    }
}
template <int R, int M, bool INV, bool CONST, int S, bool CONJ, bool HT>
__device__ __forceinline__ void levels(float2 (&a)[R], const float2* __restrict__ tw) {
    if constexpr ((1 << M) <= R) {
        level_iter<R, M, 0, INV, CONST, S, CONJ, HT>(a, tw);
        levels<R, M + 1, INV, CONST, S, CONJ, HT>(a, tw);
    }
}

// ---------------------------------------------------------------------------------------------------
// one in-register stage: radix R = 2^E DIT over the register sub-field  i = (k << SH) | g,  k in [0,R),
// for every group g in [0, 1<<SH)  (SH = 0: full radix-16 field; SH = 4-a: top field A).
//   CONST : compile-time twiddles (first stage of a direction)
//   else  : tw[e * S + g * QG] with e = 2^{m-1}-1+u; `tw` already points at the thread's Q
// ---------------------------------------------------------------------------------------------------
template <int E, int SH, int G, bool INV, bool CONST, int S, int QG, bool CONJ, bool HT, int PT>
__device__ __forceinline__ void stage_group(float2 (&v)[PT], const float2* __restrict__ tw) {
    constexpr int R = 1 << E;
    if constexpr (G < (1 << SH)) {
        float2 a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = v[(bitrev(j, E) << SH) | G];
        levels<R, 1, INV, CONST, S, CONJ, HT>(a, CONST ? nullptr : tw + G * QG);
#pragma unroll
        for (int j = 0; j < R; ++j) v[(j << SH) | G] = a[j];
        stage_group<E, SH, G + 1, INV, CONST, S, QG, CONJ, HT>(v, tw);
    }
}
template <int E, int SH, bool INV, bool CONST, int S, int QG, bool CONJ = false, bool HT = false, int PT = 16>
__device__ __forceinline__ void stage16(float2 (&v)[PT], const float2* __restrict__ tw) {
    stage_group<E, SH, 0, INV, CONST, S, QG, CONJ, HT>(v, tw);
}

// ---------------------------------------------------------------------------------------------------
// direction-specific stages (n = log2 L).  `tw` = base of ALL tables of this n in shared memory.
// ---------------------------------------------------------------------------------------------------
// first forward stage (window n-4): constant twiddles; radix 2^a if a > 0 else 16
template <int n>
__device__ __forceinline__ void fwd_first(float2 (&v)[16]) {
    constexpr int a = n % 4;
    if constexpr (a > 0) stage16<a, 4 - a, false, true, 1, 0>(v, nullptr);
    else                 stage16<4, 0, false, true, 1, 0>(v, nullptr);
}
// forward table stage of field F_f (window 4f).  The table is indexed by the thread's `hi` (position
// order, so consecutive threads read consecutive entries); k_setup_tables stores W for Q = fwd_q(hi).
template <int n, int f>
__device__ __forceinline__ void fwd_field(float2 (&v)[16], const float2* tw, int tl) {
    constexpr TwLayout lay = make_layout(n);
    constexpr int a = n % 4, nf = n / 4;
    constexpr int S = (1 << a) * ipow16(nf - 1 - f);
    static_assert(lay.fwd[f] >= 0, "field has no forward table");
    stage16<4, 0, false, false, S, 0>(v, tw + lay.fwd[f] + (tl >> (4 * f)));
}
// first inverse stage: field F_0, constant twiddles
__device__ __forceinline__ void inv_first(float2 (&v)[16]) { stage16<4, 0, true, true, 1, 0>(v, nullptr); }
// inverse table stage of field F_f, f >= 1 (window 4f): S' = 16^f, Q' = low bits of the position
template <int n, int f>
__device__ __forceinline__ void inv_field(float2 (&v)[16], const float2* tw, int tl) {
    constexpr TwLayout lay = make_layout(n);
    constexpr int S = ipow16(f);
    static_assert(lay.inv[f] >= 0, "field has no inverse table");
    stage16<4, 0, true, false, S, 0>(v, tw + lay.inv[f] + (tl & (S - 1)));
}
// inverse stage of the top field A (window n-4): radix 2^a, S' = L / 2^a, Q'(g) = tl + g * L/16
template <int n>
__device__ __forceinline__ void inv_top(float2 (&v)[16], const float2* tw, int tl) {
    constexpr TwLayout lay = make_layout(n);
    constexpr int a = n % 4;
    if constexpr (a > 0) {
        constexpr int S = (1 << n) >> a;
        stage16<a, 4 - a, true, false, S, (1 << n) / 16>(v, tw + lay.invA + tl);
    }
}

// ---------------------------------------------------------------------------------------------------
// whole-line drivers.  On entry to fwd_line the registers hold window n-4 (positions tl + (L/16) i).
// On exit they hold the F_0 window (positions 16 tl + i) and have NOT been stored.
// inv_line: entry = F_0 window in registers; exit = window n-4 in registers, not stored.
// `sync` is called between stages (a CTA barrier).
// ---------------------------------------------------------------------------------------------------
template <int n, class LAY, class Sync>
__device__ __forceinline__ void fwd_line(float2 (&v)[16], float2* line, const float2* tw, int tl, Sync sync) {
    constexpr int a = n % 4, nf = n / 4;
    constexpr int fstart = (a > 0) ? nf - 1 : nf - 2;  // first table field
    fwd_first<n>(v);
    sts16<LAY, n - 4>(v, line + LAY::base(thread_part(tl, n - 4)));
    sync();
    if constexpr (fstart >= 2) {
        float2* b = line + LAY::base(thread_part(tl, 8));
        lds16<LAY, 8>(v, b); fwd_field<n, 2>(v, tw, tl); sts16<LAY, 8>(v, b);
        sync();
    }
    if constexpr (fstart >= 1) {
        float2* b = line + LAY::base(thread_part(tl, 4));
        lds16<LAY, 4>(v, b); fwd_field<n, 1>(v, tw, tl); sts16<LAY, 4>(v, b);
        sync();
    }
    lds16<LAY, 0>(v, line + LAY::base(thread_part(tl, 0)));
    fwd_field<n, 0>(v, tw, tl);
}

template <int n, class LAY, class Sync>
__device__ __forceinline__ void inv_line(float2 (&v)[16], float2* line, const float2* tw, int tl, Sync sync) {
    constexpr int a = n % 4, nf = n / 4;
    inv_first(v);
    sts16<LAY, 0>(v, line + LAY::base(thread_part(tl, 0)));
    sync();
    if constexpr (nf >= 2) {
        float2* b = line + LAY::base(thread_part(tl, 4));
        lds16<LAY, 4>(v, b); inv_field<n, 1>(v, tw, tl);
        if constexpr (!(a == 0 && nf == 2)) { sts16<LAY, 4>(v, b); sync(); }
    }
    if constexpr (nf >= 3) {
        float2* b = line + LAY::base(thread_part(tl, 8));
        lds16<LAY, 8>(v, b); inv_field<n, 2>(v, tw, tl);
        if constexpr (!(a == 0 && nf == 3)) { sts16<LAY, 8>(v, b); sync(); }
    }
    if constexpr (a > 0) {
        lds16<LAY, n - 4>(v, line + LAY::base(thread_part(tl, n - 4)));
        inv_top<n>(v, tw, tl);
    }
}

// ---------------------------------------------------------------------------------------------------
// 32 points per thread (used for L = 1024 = 32 x 32: ONE exchange per FFT, a row is private to one warp)
// ---------------------------------------------------------------------------------------------------
struct RowLayout32 {   // one pad element after every 32: a thread's 32 contiguous positions are 33 from its neighbour's
    __host__ __device__ static constexpr int phys(int pos) { return pos + (pos >> 5); }
    __host__ __device__ static constexpr int line_elems(int L) { return L + (L >> 5); }
    __host__ __device__ static constexpr int off(int i, int w0) { return phys(i << w0); }
};
template <int CC>
struct ColLayout32 {   // [padded row][CC columns]
    __host__ __device__ static constexpr int rows(int L) { return L + (L >> 5); }
    __host__ __device__ static constexpr int off(int i, int w0) { return RowLayout32::phys(i << w0) * CC; }
};
// radix-32 stages of the 1024-point transform.  `tw` = forward HALF table [16][32]: entry (off(m) + u, Q) =
// W_{32 2^m}^{Q + 32 u}, u < max(1, 2^{m-2}) (see bfly); the inverse reads the same table conjugated.
__device__ __forceinline__ void fwd32_first(float2 (&v)[32]) { stage16<5, 0, false, true, 1, 0>(v, nullptr); }
__device__ __forceinline__ void inv32_first(float2 (&v)[32]) { stage16<5, 0, true, true, 1, 0>(v, nullptr); }
__device__ __forceinline__ void fwd32_table(float2 (&v)[32], const float2* tw_q) { stage16<5, 0, false, false, 32, 0, false, true>(v, tw_q); }
__device__ __forceinline__ void inv32_table(float2 (&v)[32], const float2* tw_q) { stage16<5, 0, true, false, 32, 0, true, true>(v, tw_q); }

}  // namespace asmb

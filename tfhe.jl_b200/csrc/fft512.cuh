// fft512.cuh — negacyclic transform of degree-1024 real polynomials for sm_100a.
//
// Replaces forward_transform / inverse_transform of the reference (polynomials.jl:106-132), which
// call FFTW on a folded N/2-point complex-double buffer.  Same mathematics, own dataflow:
//
//   p(X) mod (X^1024 + 1), real coefficients  ->  z_j = p_j - i*p_{j+512}   (= p mod X^512 + i)
//   Z_k = z(zeta^(4k+1)),  zeta = exp(-i*pi/1024),  k = 0..511              (roots of X^512 = -i)
//
// A 64-thread group owns one transform; every thread keeps 8 complex points in registers and the
// transform is three radix-8 passes with two exchanges through shared memory:
//
//   pass 1: thread t holds j = t + 64m (m = 0..7).  The twist exp(-i*pi*j/1024) is factored as
//           exp(-i*pi*t/1024) * exp(-i*pi*m/16): the second factor is a compile-time constant applied
//           before the radix-8 butterfly, the first is merged with the pass-1 twiddle into
//           T_q(t) = E(t*(4q+1)), E(x) = exp(-i*pi*x/1024).
//   exchange 1 (X1): (t1,t2 | q) -> thread t1 + 8q, register t2          [t = t1 + 8*t2]
//   pass 2: radix-8 over t2 -> q2, twiddle W64^(t1*q2) = E(32*t1*q2)
//   exchange 2 (X2): (t1,q | q2) -> thread q2 + 8q, register t1          [8x8 tiles padded to 9x8: conflict-free]
//   pass 3: radix-8 over t1 -> q3
//
// Output: thread v = q2 + 8q, register q3 holds frequency k = q + 8*q2 + 64*q3.  The transform-domain
// layout of every stored spectrum is therefore [q3][v] (8 x 64 complex), which makes the pointwise
// multiply-accumulate a perfectly coalesced 16-byte load per thread, and the inverse transform is the
// exact mirror (no reordering pass anywhere).
//
// Both exchanges use 16-byte (double2) shared-memory accesses; a quarter-warp always touches 8
// distinct 16-byte bank groups, so all LDS.128/STS.128 are conflict-free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

constexpr int kN = 1024;          // polynomial degree (all parameter sets of the reference use 1024)
constexpr int kHalf = 512;        // complex points per transform
constexpr int kGroup = 64;        // threads per transform
constexpr int kSpectrum = 512;    // double2 per stored spectrum
constexpr int kX2Elems = 576;     // double2 in the second exchange buffer: 8 rows of 8x8 tiles padded to 9x8 (conflict-free
                                  // transposes with compile-time offsets instead of an XOR swizzle)

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y)));
}
// acc += a * b
__device__ __forceinline__ void cmac(double2& acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
template <bool INV> __device__ __forceinline__ double2 tw(double2 a, double2 w) { return INV ? cmulc(a, w) : cmul(a, w); }
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ double2 rot90(double2 a) {
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

// 64-thread group barrier (barrier 0 is left to __syncthreads)
__device__ __forceinline__ void group_sync(int bar_id) {
    asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
}

// 8-point DFT in registers, natural order in and out.  Forward kernel exp(-2*pi*i*m*q/8); INV conjugates.
template <bool INV> __device__ __forceinline__ void dft8(double2 (&a)[8]) {
    constexpr double h = 0.70710678118654752440;
    double2 s0 = cadd(a[0], a[4]), d0 = csub(a[0], a[4]);
    double2 s1 = cadd(a[1], a[5]), d1 = csub(a[1], a[5]);
    double2 s2 = cadd(a[2], a[6]), d2 = csub(a[2], a[6]);
    double2 s3 = cadd(a[3], a[7]), d3 = csub(a[3], a[7]);
    // d_m *= W8^m.  W8^1 and W8^3 are (+-1 +- i)/sqrt(2): the factor h is not applied here but folded into the
    // last butterfly as an FMA (a = f +- h*g), which saves the four multiplications.
    if (!INV) {
        d1 = make_double2(d1.x + d1.y, d1.y - d1.x);
        d3 = make_double2(d3.y - d3.x, -(d3.x + d3.y));
    } else {
        d1 = make_double2(d1.x - d1.y, d1.x + d1.y);
        d3 = make_double2(-(d3.x + d3.y), d3.x - d3.y);
    }
    d2 = rot90<INV>(d2);
    double2 e0 = cadd(s0, s2), e1 = csub(s0, s2), o0 = cadd(s1, s3), o1 = rot90<INV>(csub(s1, s3));
    double2 f0 = cadd(d0, d2), f1 = csub(d0, d2), g0 = cadd(d1, d3), g1 = rot90<INV>(csub(d1, d3));   // g0, g1 unscaled
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0); a[2] = cadd(e1, o1); a[6] = csub(e1, o1);
    a[1] = make_double2(fma(h, g0.x, f0.x), fma(h, g0.y, f0.y)); a[5] = make_double2(fma(-h, g0.x, f0.x), fma(-h, g0.y, f0.y));
    a[3] = make_double2(fma(h, g1.x, f1.x), fma(h, g1.y, f1.y)); a[7] = make_double2(fma(-h, g1.x, f1.x), fma(-h, g1.y, f1.y));
}

// exp(-i*pi*m/16), m = 0..7: the thread-independent part of the twist
__device__ __constant__ double kTwistRe[8] = {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452,
                                              0.7071067811865476, 0.5555702330196023, 0.38268343236508984,
                                              0.19509032201612833};
__device__ __constant__ double kTwistIm[8] = {-0.0, -0.19509032201612825, -0.3826834323650898, -0.5555702330196022,
                                              -0.7071067811865475, -0.8314696123025452, -0.9238795325112867,
                                              -0.9807852804032304};

// Per-thread twiddle bases, loaded once per kernel from the table E[x] = exp(-i*pi*x/1024), x < 2048.
struct Twiddles {
    double2 e1, s1, s2, s4;  // E(t), E(4t), E(8t), E(16t)     -> T_q = E(t*(4q+1))
    double2 v1, v2, v4;      // E(32*t1), E(64*t1), E(128*t1)  -> W64^(t1*q2), t1 = t & 7
    __device__ __forceinline__ void load(const double2* __restrict__ E, int t) {
        e1 = E[t]; s1 = E[4 * t]; s2 = E[8 * t]; s4 = E[16 * t];
        int t1 = t & 7;
        v1 = E[32 * t1]; v2 = E[64 * t1]; v4 = E[128 * t1];
    }
};

// All 15 twiddles of a thread, read once from the table (every entry correctly rounded from long double) instead of
// being re-derived by 11 complex multiplications in every transform: 44 fewer FP64 instructions per transform and
// thread for 32 more registers.  Used by the kernels whose register budget allows it (blind_rotate.cuh, TM == 3).
struct TwiddlesFull {
    double2 T[8];   // T[q]  = E(t*(4q+1))   pass-1 twiddle merged with the per-thread part of the twist
    double2 V[8];   // V[q2] = E(32*t1*q2) = W64^(t1*q2), t1 = t & 7;  V[0] = 1 is never used
    __device__ __forceinline__ void load(const double2* __restrict__ E, int t) {
#pragma unroll
        for (int q = 0; q < 8; q++) T[q] = E[t * (4 * q + 1)];          // <= 63*29 = 1827 < 2048
        const int t1 = t & 7;
        V[0] = make_double2(1.0, 0.0);
#pragma unroll
        for (int q2 = 1; q2 < 8; q2++) V[q2] = E[32 * t1 * q2];         // <= 32*49 = 1568
    }
};

__device__ __forceinline__ void pass1_twiddles(const Twiddles& w, double2 (&T)[8]) {
    T[0] = w.e1; T[1] = cmul(w.e1, w.s1); T[2] = cmul(w.e1, w.s2); T[3] = cmul(T[1], w.s2); T[4] = cmul(w.e1, w.s4);
    T[5] = cmul(T[1], w.s4); T[6] = cmul(T[2], w.s4); T[7] = cmul(T[3], w.s4);
}
__device__ __forceinline__ void pass1_twiddles(const TwiddlesFull& w, double2 (&T)[8]) {
#pragma unroll
    for (int q = 0; q < 8; q++) T[q] = w.T[q];
}
__device__ __forceinline__ void pass2_twiddles(const Twiddles& w, double2 (&V)[8]) {
    V[1] = w.v1; V[2] = w.v2; V[3] = cmul(w.v1, w.v2); V[4] = w.v4; V[5] = cmul(w.v1, w.v4); V[6] = cmul(w.v2, w.v4);
    V[7] = cmul(V[3], w.v4);
}
__device__ __forceinline__ void pass2_twiddles(const TwiddlesFull& w, double2 (&V)[8]) {
#pragma unroll
    for (int q = 1; q < 8; q++) V[q] = w.V[q];
}

// Forward transform.  In: a[m] = z_{t+64m} (untwisted).  Out: a[q3] = Z_{q + 8*q2 + 64*q3}, v = q2 + 8q = t.
// X1, X2: two 512-element double2 scratch buffers private to the 64-thread group.
struct NoPrefetch { __device__ __forceinline__ void operator()() const {} };

// `prefetch` runs right after the second barrier, before the X2 loads are issued: callers use it to put
// further shared-memory loads (the key chunk of the multiply-accumulate) in flight behind the same wait.
// SYNC == 0: both exchanges are fenced by the 64-thread group barrier; X1 may be the same buffer in every call.
// SYNC == 1: the second exchange stays inside 8-thread tiles (threads 8*hi .. 8*hi+7, all in one warp), so it only
//            needs __syncwarp(); the one group barrier left per transform sits between the X1 stores and loads.
//            The caller must then ALTERNATE between two X1 buffers from one transform to the next (a warp that has
//            passed the barrier of transform k may already store pass-1 results of transform k+1 while the other warp
//            of the group still loads X1 of transform k; it cannot reach transform k+2 before that warp has arrived at
//            the barrier of transform k+1, i.e. has consumed those loads).
// (An in-place variant of the second exchange — XOR-swizzled inside the tile's own 1 KB block of X1, no X2 buffer,
// two more ring stages — was bit-exact and 2 % slower: profiles/r2/k3_variants.md.)
template <int SYNC, class W, class F>
__device__ __forceinline__ void fft512_forward_t(double2 (&a)[8], const W& w, double2* X1, double2* X2, int t,
                                                 int bar_id, F&& prefetch) {
#pragma unroll
    for (int m = 1; m < 8; m++) a[m] = cmul(a[m], make_double2(kTwistRe[m], kTwistIm[m]));
    dft8<false>(a);
    {
        double2 T[8];
        pass1_twiddles(w, T);
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = cmul(a[q], T[q]);
    }
#pragma unroll
    for (int q = 0; q < 8; q++) X1[q * 64 + t] = a[q];
    group_sync(bar_id);
    const int lo = t & 7, hi = t >> 3;
#pragma unroll
    for (int t2 = 0; t2 < 8; t2++) a[t2] = X1[hi * 64 + lo + 8 * t2];
    dft8<false>(a);
    {
        double2 V[8];
        pass2_twiddles(w, V);
#pragma unroll
        for (int q = 1; q < 8; q++) a[q] = cmul(a[q], V[q]);
    }
    if (SYNC == 1) __syncwarp();   // the tile's loads of the previous transform's X2 are done
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) X2[hi * 72 + q2 * 9 + lo] = a[q2];
    if (SYNC == 1) __syncwarp(); else group_sync(bar_id);
    prefetch();
#pragma unroll
    for (int t1 = 0; t1 < 8; t1++) a[t1] = X2[hi * 72 + lo * 9 + t1];
    dft8<false>(a);
}
template <class W, class F = NoPrefetch>
__device__ __forceinline__ void fft512_forward(double2 (&a)[8], const W& w, double2* X1, double2* X2, int t, int bar_id,
                                               F&& prefetch = F()) {
    fft512_forward_t<0>(a, w, X1, X2, t, bar_id, prefetch);
}

// Inverse transform, the mirror of fft512_forward.  In: a[q3] spectrum at thread v.  Out: a[m] = z_{t+64m}
// scaled by 1/512 and untwisted, i.e. the folded coefficients (p_j - i*p_{j+512}).
template <int SYNC, class W>
__device__ __forceinline__ void fft512_inverse_t(double2 (&a)[8], const W& w, double2* X1, double2* X2, int t, int bar_id) {
    const int lo = t & 7, hi = t >> 3;
    dft8<true>(a);   // q3 -> t1
    {
        double2 V[8];
        pass2_twiddles(w, V);
#pragma unroll
        for (int q = 1; q < 8; q++) a[q] = cmulc(a[q], V[q]);
    }
    if (SYNC == 1) __syncwarp();
#pragma unroll
    for (int t1 = 0; t1 < 8; t1++) X2[hi * 72 + lo * 9 + t1] = a[t1];
    if (SYNC == 1) __syncwarp(); else group_sync(bar_id);
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) a[q2] = X2[hi * 72 + q2 * 9 + lo];
    dft8<true>(a);   // q2 -> t2
#pragma unroll
    for (int t2 = 0; t2 < 8; t2++) X1[hi * 64 + lo + 8 * t2] = a[t2];
    group_sync(bar_id);
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = X1[q * 64 + t];
    {
        double2 T[8];
        pass1_twiddles(w, T);
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = cmulc(a[q], T[q]);
    }
    dft8<true>(a);   // q -> m
    constexpr double sc = 1.0 / 512.0;
    a[0] = make_double2(a[0].x * sc, a[0].y * sc);
#pragma unroll
    for (int m = 1; m < 8; m++) a[m] = cmulc(a[m], make_double2(kTwistRe[m] * sc, kTwistIm[m] * sc));
}
template <class W>
__device__ __forceinline__ void fft512_inverse(double2 (&a)[8], const W& w, double2* X1, double2* X2, int t, int bar_id) {
    fft512_inverse_t<0>(a, w, X1, X2, t, bar_id);
}

// Two transforms per thread in one instruction stream (A through X1a, B through X1b, the tile-local second exchange of
// both through the one X2 buffer, one after the other).  The arithmetic of one transform is independent of the other's
// exchanges, so the scheduler can cover the shared-memory latency of A with butterflies of B inside a single warp.
// Both X1 buffers are in use in every call, so the call opens with a group barrier (every thread has finished the X1
// loads of the previous call) — two barriers per pair = one per transform, as in the SYNC == 1 single transforms.
template <class W>
__device__ __forceinline__ void fft512_forward_dual(double2 (&a)[8], double2 (&b)[8], const W& w, double2* X1a, double2* X1b,
                                                    double2* X2, int t, int bar_id) {
#pragma unroll
    for (int m = 1; m < 8; m++) {
        a[m] = cmul(a[m], make_double2(kTwistRe[m], kTwistIm[m]));
        b[m] = cmul(b[m], make_double2(kTwistRe[m], kTwistIm[m]));
    }
    dft8<false>(a); dft8<false>(b);
    {
        double2 T[8];
        pass1_twiddles(w, T);
#pragma unroll
        for (int q = 0; q < 8; q++) { a[q] = cmul(a[q], T[q]); b[q] = cmul(b[q], T[q]); }
    }
    group_sync(bar_id);
#pragma unroll
    for (int q = 0; q < 8; q++) { X1a[q * 64 + t] = a[q]; X1b[q * 64 + t] = b[q]; }
    group_sync(bar_id);
    const int lo = t & 7, hi = t >> 3;
#pragma unroll
    for (int t2 = 0; t2 < 8; t2++) { a[t2] = X1a[hi * 64 + lo + 8 * t2]; b[t2] = X1b[hi * 64 + lo + 8 * t2]; }
    double2 V[8];
    pass2_twiddles(w, V);
    dft8<false>(a);
#pragma unroll
    for (int q = 1; q < 8; q++) a[q] = cmul(a[q], V[q]);
    __syncwarp();
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) X2[hi * 72 + q2 * 9 + lo] = a[q2];
    __syncwarp();
#pragma unroll
    for (int t1 = 0; t1 < 8; t1++) a[t1] = X2[hi * 72 + lo * 9 + t1];
    dft8<false>(b);
#pragma unroll
    for (int q = 1; q < 8; q++) b[q] = cmul(b[q], V[q]);
    __syncwarp();
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) X2[hi * 72 + q2 * 9 + lo] = b[q2];
    __syncwarp();
#pragma unroll
    for (int t1 = 0; t1 < 8; t1++) b[t1] = X2[hi * 72 + lo * 9 + t1];
    dft8<false>(a); dft8<false>(b);
}

template <class W>
__device__ __forceinline__ void fft512_inverse_dual(double2 (&a)[8], double2 (&b)[8], const W& w, double2* X1a, double2* X1b,
                                                    double2* X2, int t, int bar_id) {
    const int lo = t & 7, hi = t >> 3;
    double2 V[8];
    pass2_twiddles(w, V);
    dft8<true>(a);
#pragma unroll
    for (int q = 1; q < 8; q++) a[q] = cmulc(a[q], V[q]);
    __syncwarp();
#pragma unroll
    for (int t1 = 0; t1 < 8; t1++) X2[hi * 72 + lo * 9 + t1] = a[t1];
    __syncwarp();
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) a[q2] = X2[hi * 72 + q2 * 9 + lo];
    dft8<true>(b);
#pragma unroll
    for (int q = 1; q < 8; q++) b[q] = cmulc(b[q], V[q]);
    __syncwarp();
#pragma unroll
    for (int t1 = 0; t1 < 8; t1++) X2[hi * 72 + lo * 9 + t1] = b[t1];
    __syncwarp();
#pragma unroll
    for (int q2 = 0; q2 < 8; q2++) b[q2] = X2[hi * 72 + q2 * 9 + lo];
    dft8<true>(a); dft8<true>(b);
    group_sync(bar_id);
#pragma unroll
    for (int t2 = 0; t2 < 8; t2++) { X1a[hi * 64 + lo + 8 * t2] = a[t2]; X1b[hi * 64 + lo + 8 * t2] = b[t2]; }
    group_sync(bar_id);
#pragma unroll
    for (int q = 0; q < 8; q++) { a[q] = X1a[q * 64 + t]; b[q] = X1b[q * 64 + t]; }
    {
        double2 T[8];
        pass1_twiddles(w, T);
#pragma unroll
        for (int q = 0; q < 8; q++) { a[q] = cmulc(a[q], T[q]); b[q] = cmulc(b[q], T[q]); }
    }
    dft8<true>(a); dft8<true>(b);
    constexpr double sc = 1.0 / 512.0;
    a[0] = make_double2(a[0].x * sc, a[0].y * sc);
    b[0] = make_double2(b[0].x * sc, b[0].y * sc);
#pragma unroll
    for (int m = 1; m < 8; m++) {
        a[m] = cmulc(a[m], make_double2(kTwistRe[m] * sc, kTwistIm[m] * sc));
        b[m] = cmulc(b[m], make_double2(kTwistRe[m] * sc, kTwistIm[m] * sc));
    }
}

// round-to-nearest-even to a 64-bit integer, keep the low 32 bits (polynomials.jl:115-116)
__device__ __forceinline__ uint32_t round_to_u32(double x) { return (uint32_t)(unsigned long long)__double2ll_rn(x); }

// The same rounding for |x| < 2^51 with one FP64 add instead of F2I.S64 (a quarter-rate conversion: 14/clk/SM
// measured, profiles/r1/microbench.json): x + 1.5*2^52 is rounded to nearest-even by the adder and the low
// mantissa word is round(x) mod 2^32.  PROVEN == true (split transform) bounds |x| by 2^41, so it is exact;
// the unsplit transform has no such bound and keeps the conversion instruction.
template <bool PROVEN> __device__ __forceinline__ uint32_t round_to_u32_fast(double x) {
    if (PROVEN) return (uint32_t)__double2loint(x + 6755399441055744.0);
    return round_to_u32(x);
}

// digit value (an integer in [0, 2^21)) minus `half`, as a double, without I2F: 2^52 + u has u in its low
// mantissa word, and subtracting (2^52 + half) is exact.
__device__ __forceinline__ double small_uint_minus_half_to_double(uint32_t u, int half) {
    return __hiloint2double(0x43300000, (int)u) - (4503599627370496.0 + (double)half);
}

}  // namespace tfhe_b200

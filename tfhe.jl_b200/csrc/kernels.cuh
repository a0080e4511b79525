// kernels.cuh — sm_100a kernels of the single-key gate-bootstrapping path.
//
//   K6  bk_transform_kernel     forward_transform.(bk)            bootstrap.jl:12, tgsw.jl:120-121
//   K1  polymul_kernel          transformed_mul                   polynomials.jl:142-144
//   K2  extern_product_step     decompose + tgsw_extern_mul       tgsw.jl:99-129
//   K3  blind_rotate_kernel     gate prologue + modswitch + blind_rotate + extract
//                                                                 gates.jl:15-153, bootstrap.jl:19-82, tlwe.jl:55-59
//   K4  keyswitch_kernel        keyswitch                         keyswitch.jl:45-80
//
// Exactness.  All products are computed with the complex-double transform of fft512.cuh.  With
// NP == 2 every torus operand (a bootstrapping-key polynomial) is split into two signed 16-bit
// pieces before it is transformed, so every real number that is rounded back to an integer is
// bounded by (k+1)*l*N*(Bg/2)*2^15 <= 2^36 and the floating-point error is provably < 2^-10
// (DESIGN.md §Exactness); the result is reassembled as lo + (hi << 16) mod 2^32 and is therefore the
// exact integer negacyclic convolution.  NP == 1 is the reference's own regime (one 32-bit piece,
// polynomials.jl:138-140): exact in practice, but without a proof.
#pragma once
#include "fft512.cuh"

namespace tfhe_b200 {

// ------------------------------------------------------------------------------------------------
// small integer helpers (torus arithmetic is uint32 wrap-around)

// coefficient x of X^s * p mod (X^1024 + 1), s in [0, 2048)   (DarkIntegers mul_by_monomial)
__device__ __forceinline__ int32_t rot_coeff(const int32_t* __restrict__ p, int x, int s) {
    int y = (x - s) & 2047;
    int32_t v = p[y & 1023];
    return (y & 1024) ? (int32_t)(0u - (uint32_t)v) : v;
}

// decompose (tgsw.jl:99-117), digit r (0-based): bits [32-(r+1)*BGBIT, 32-r*BGBIT) of x + offset, minus Bg/2
template <int L, int BGBIT> __host__ __device__ constexpr uint32_t decomp_offset() {
    uint32_t s = 0;
    for (int r = 1; r <= L; r++) s += 1u << (32 - r * BGBIT);
    return s * (1u << (BGBIT - 1));
}
template <int BGBIT> __device__ __forceinline__ int32_t digit(uint32_t x_plus_offset, int r) {
    return (int32_t)((x_plus_offset >> (32 - (r + 1) * BGBIT)) & ((1u << BGBIT) - 1)) - (1 << (BGBIT - 1));
}

// the same digit as a double (no integer-to-float conversion instruction, see fft512.cuh)
template <int BGBIT> __device__ __forceinline__ double digit_f64(uint32_t x_plus_offset, int r) {
    return small_uint_minus_half_to_double((x_plus_offset >> (32 - (r + 1) * BGBIT)) & ((1u << BGBIT) - 1), 1 << (BGBIT - 1));
}

// decode_message(x, 2N) (numeric-functions.jl:31-34) for 2N = 2048: (x + 2^20) >> 21, arithmetic
__device__ __forceinline__ int32_t modswitch2048(int32_t x) { return (int32_t)((uint32_t)x + (1u << 20)) >> 21; }

// split a torus word into signed 16-bit pieces: x = hi * 2^16 + lo, lo in [-2^15, 2^15)
__device__ __forceinline__ void split16(int32_t x, int32_t& lo, int32_t& hi) {
    lo = (int32_t)(int16_t)(x & 0xffff);
    hi = (int32_t)(((int64_t)x - lo) >> 16);
}

// ------------------------------------------------------------------------------------------------
// K6: coefficient-domain polynomials -> stored spectra.  One 64-thread CTA per polynomial.
// out layout: [poly][piece][q3*64 + v]
template <int NP>
__global__ void __launch_bounds__(64) bk_transform_kernel(const int32_t* __restrict__ polys,
                                                          double2* __restrict__ out,
                                                          const double2* __restrict__ E) {
    __shared__ double2 X1[512];
    __shared__ double2 X2[kX2Elems];
    const int t = threadIdx.x;
    Twiddles w; w.load(E, t);
    const int32_t* p = polys + (size_t)blockIdx.x * kN;
#pragma unroll 1
    for (int piece = 0; piece < NP; piece++) {
        double2 a[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int32_t c0 = p[t + 64 * m], c1 = p[t + 64 * m + 512];
            if (NP == 2) {
                int32_t l0, h0, l1, h1;
                split16(c0, l0, h0); split16(c1, l1, h1);
                c0 = piece ? h0 : l0; c1 = piece ? h1 : l1;
            }
            a[m] = make_double2((double)c0, -(double)c1);
        }
        fft512_forward(a, w, X1, X2, t, 0);
        double2* o = out + ((size_t)blockIdx.x * NP + piece) * kSpectrum;
#pragma unroll
        for (int q3 = 0; q3 < 8; q3++) o[q3 * 64 + t] = a[q3];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// K2: one external product on an accumulator held in shared memory.
//   temp_c = ROTSUB ? X^abar * acc_c - acc_c : acc_c                      (bootstrap.jl:21 / tgsw.jl:126)
//   res_c' = sum_{r,c} digit_r(temp_c) (*) BK[r][c][c']                    (tgsw.jl:126-128)
//   acc_c' = ACCUM ? acc_c' + res_c' : res_c'                              (bootstrap.jl:22)
// bk_row points at [r][c][c'][piece][q3][v] for this key element.  The 64 threads of the group all call it.
//
// Output spectra: (k+1)*NP accumulators of 8 complex points per thread, all in registers (64 registers for
// NP == 1, 128 for NP == 2).  That only works because K3 runs 8 warps per SM (2 per sub-partition, up to 255
// registers per thread): the 168-register v1 build spilled 3.0e9 local ld/st per 4096 gates and wrote 15.9 GB
// to DRAM, and keeping half of them in shared memory (v2/v3a) cost 23 % more shared-memory wavefronts.
//
// Where the bootstrapping-key spectra of one key element come from.  A "chunk" is 2 spectra = 16 KB:
// chunk (r, c, half) = index (r*2 + c)*NP + half of the row; for NP == 1 it holds output components
// c' = 0,1, for NP == 2 it holds the two 16-bit pieces of output component c' = half.
struct BkFromGlobal {   // coalesced 16-byte read-only loads straight from L2 (stand-alone K2 kernel, tests)
    const double2* row;
    __device__ __forceinline__ const double2* acquire(int chunk) { return row + (size_t)chunk * 2 * kSpectrum; }
    __device__ __forceinline__ void release() {}
    static __device__ __forceinline__ double2 load(const double2* p) { return __ldg(p); }
};

template <int L, int BGBIT, int NP, bool ROTSUB, bool ACCUM, class BK>
__device__ __forceinline__ void extern_product_step(int32_t* acc, int abar, BK& bk,
                                                    const Twiddles& w, double2* X1, double2* X2, int t, int bar_id) {
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
    double2 o[2][NP][8];
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++)
#pragma unroll
        for (int pc = 0; pc < NP; pc++)
#pragma unroll
            for (int q = 0; q < 8; q++) o[c2][pc][q] = make_double2(0.0, 0.0);

#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        const int32_t* p = acc + c * kN;
        uint32_t tl[8], th[8];   // temp_c at j and j+512, offset already added
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            if (ROTSUB) {
                tl[m] = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;
                th[m] = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
            } else {
                tl[m] = (uint32_t)p[j] + offset;
                th[m] = (uint32_t)p[j + 512] + offset;
            }
        }
#pragma unroll 1
        for (int r = 0; r < L; r++) {
            double2 a[8];
#pragma unroll
            for (int m = 0; m < 8; m++)
                a[m] = make_double2(digit_f64<BGBIT>(tl[m], r), -digit_f64<BGBIT>(th[m], r));
            fft512_forward(a, w, X1, X2, t, bar_id);
#pragma unroll
            for (int half = 0; half < NP; half++) {
                // chunk (r, c, half): both output components (NP == 1) or the two pieces of component `half`
                const double2* b = bk.acquire((r * 2 + c) * NP + half) + t;
#pragma unroll
                for (int sp = 0; sp < 2; sp++)
#pragma unroll
                    for (int q = 0; q < 8; q++)
                        cmac(NP == 1 ? o[sp][0][q] : o[half][sp % NP][q], a[q], BK::load(b + (sp * 8 + q) * 64));
                bk.release();
            }
        }
    }
    // every thread has finished reading acc and X2 (last forward) before anyone overwrites them
    group_sync(bar_id);
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++) {
        double2 (&oc)[NP][8] = o[c2];
        uint32_t rl[8], rh[8];
#pragma unroll
        for (int pc = 0; pc < NP; pc++) {
            fft512_inverse(oc[pc], w, X1, X2, t, bar_id);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                uint32_t vl = round_to_u32_fast<NP == 2>(oc[pc][m].x), vh = round_to_u32_fast<NP == 2>(-oc[pc][m].y);
                if (pc == 0) { rl[m] = vl; rh[m] = vh; }
                else { rl[m] += vl << 16; rh[m] += vh << 16; }
            }
        }
        int32_t* p = acc + c2 * kN;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            if (ACCUM) { p[j] = (int32_t)((uint32_t)p[j] + rl[m]); p[j + 512] = (int32_t)((uint32_t)p[j + 512] + rh[m]); }
            else { p[j] = (int32_t)rl[m]; p[j + 512] = (int32_t)rh[m]; }
        }
    }
    group_sync(bar_id);
}

// K2 as a stand-alone batch kernel (parity tests of tgsw_extern_mul): one group per product
template <int L, int BGBIT, int NP>
__global__ void __launch_bounds__(64) extern_product_kernel(const double2* __restrict__ bk_fft,
                                                            const double2* __restrict__ E,
                                                            const int32_t* __restrict__ acc_in,
                                                            const int32_t* __restrict__ bk_index,
                                                            int32_t* __restrict__ out) {
    __shared__ double2 X1[512];
    __shared__ double2 X2[kX2Elems];
    __shared__ int32_t acc[2 * kN];
    const int t = threadIdx.x;
    Twiddles w; w.load(E, t);
    const size_t g = blockIdx.x;
    for (int x = t; x < 2 * kN; x += 64) acc[x] = acc_in[g * 2 * kN + x];
    __syncthreads();
    const size_t row = (size_t)L * 2 * 2 * NP * kSpectrum;
    BkFromGlobal bk{bk_fft + (size_t)bk_index[g] * row};
    extern_product_step<L, BGBIT, NP, false, false>(acc, 0, bk, w, X1, X2, t, 0);
    for (int x = t; x < 2 * kN; x += 64) out[g * 2 * kN + x] = acc[x];
}

// ------------------------------------------------------------------------------------------------
// K1: exact negacyclic product of two arbitrary int32 polynomials mod 2^32.
// x = xh*2^16 + xl, y = yh*2^16 + yl (signed 16-bit pieces):  x*y = xl*yl + 2^16 (xh*yl + xl*yh)  (mod 2^32)
// Rounded magnitudes <= 2^41, error bound < 0.03.
// x_stride: distance in words between consecutive x operands (kN; 0 = the same x for every product, e.g. the TLWE key);
// y_stride / out_stride likewise (k * kN when the operands are one mask polynomial of TLWE samples with k of them)
__global__ void __launch_bounds__(64) polymul_kernel(const int32_t* __restrict__ xs, const int32_t* __restrict__ ys,
                                                     int32_t* __restrict__ out, const double2* __restrict__ E,
                                                     size_t x_stride = kN, size_t y_stride = kN, size_t out_stride = kN) {
    __shared__ double2 X1[512];
    __shared__ double2 X2[kX2Elems];
    __shared__ double2 SX[2][512];   // spectra of xl, xh
    const int t = threadIdx.x;
    Twiddles w; w.load(E, t);
    const int32_t* x = xs + (size_t)blockIdx.x * x_stride;
    const int32_t* y = ys + (size_t)blockIdx.x * y_stride;
    double2 a[8];
#pragma unroll 1
    for (int pc = 0; pc < 2; pc++) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int32_t l0, h0, l1, h1;
            split16(x[t + 64 * m], l0, h0); split16(x[t + 64 * m + 512], l1, h1);
            a[m] = make_double2((double)(pc ? h0 : l0), -(double)(pc ? h1 : l1));
        }
        fft512_forward(a, w, X1, X2, t, 0);
#pragma unroll
        for (int q = 0; q < 8; q++) SX[pc][q * 64 + t] = a[q];
    }
    double2 p1[8], p2[8];
#pragma unroll 1
    for (int pc = 0; pc < 2; pc++) {
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int32_t l0, h0, l1, h1;
            split16(y[t + 64 * m], l0, h0); split16(y[t + 64 * m + 512], l1, h1);
            a[m] = make_double2((double)(pc ? h0 : l0), -(double)(pc ? h1 : l1));
        }
        fft512_forward(a, w, X1, X2, t, 0);
#pragma unroll
        for (int q = 0; q < 8; q++) {
            double2 xl = SX[0][q * 64 + t], xh = SX[1][q * 64 + t];   // own entries: no barrier needed
            if (pc == 0) { p1[q] = cmul(xl, a[q]); p2[q] = cmul(xh, a[q]); }
            else cmac(p2[q], xl, a[q]);
        }
    }
    __syncthreads();
    fft512_inverse(p1, w, X1, X2, t, 0);
    fft512_inverse(p2, w, X1, X2, t, 0);
    int32_t* o = out + (size_t)blockIdx.x * out_stride;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        o[t + 64 * m] = (int32_t)(round_to_u32_fast<true>(p1[m].x) + (round_to_u32_fast<true>(p2[m].x) << 16));
        o[t + 64 * m + 512] = (int32_t)(round_to_u32_fast<true>(-p1[m].y) + (round_to_u32_fast<true>(-p2[m].y) << 16));
    }
}

// ------------------------------------------------------------------------------------------------
// K4: LWE key switch (keyswitch.jl:45-80) as a gather-accumulate over an L2-resident table.
// ksk rows are padded to `stride` words (multiple of 32 -> 128-byte aligned rows); one CTA per
// ciphertext, stride/4 threads, each owning 4 consecutive output columns (one LDG.128 per row).
// If `out_stride_words`/`in_stride_words` differ from the natural widths the same kernel serves the
// per-party key switches of mk_keyswitch (mk_internals.jl:397-411).
struct KeyswitchArgs {
    const int32_t* ksk;      // [Nk][t][base-1][stride]
    const int32_t* in;       // ciphertext g, mask i at in[g*in_stride + in_offset + i]
    int32_t* out;            // out[g*out_stride + out_offset + c], c < n
    int32_t* out_b;          // where column n is accumulated: out_b[g*out_stride + b_offset]
    int n, Nk, t, basebit, stride;
    long long in_stride, in_offset, out_stride, out_offset, b_offset;
    int b_mode;              // 0: out_b = in_b - sum ; 1: out_b += (0 - sum)  (MK: parts added to the joint b)
    long long in_b_offset;   // position of the input b (b_mode 0)
};

// (A 16-ciphertexts-per-CTA lockstep variant that shares table rows through L1 cut L2 traffic 4x but executed
// 2.4x more instructions and was slower, 18.0 vs 15.0 ms per 16 384 ciphertexts; this version runs at the L2
// bandwidth limit, 13.8 TB/s — profiles/r1/ncu_v3_keyswitch.txt.)
__global__ void keyswitch_kernel(KeyswitchArgs A, unsigned long long count) {
    extern __shared__ int32_t s_a[];
    const size_t g = blockIdx.x;
    if (g >= count) return;
    const int32_t* in = A.in + g * A.in_stride + A.in_offset;
    const uint32_t prec = 1u << (32 - (1 + A.basebit * A.t));   // keyswitch.jl:58
    for (int i = threadIdx.x; i < A.Nk; i += blockDim.x) s_a[i] = (int32_t)((uint32_t)in[i] + prec);
    __syncthreads();
    const int base1 = (1 << A.basebit) - 1;
    const uint32_t mask = (uint32_t)base1;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint4* rows = reinterpret_cast<const uint4*>(A.ksk) + threadIdx.x;
    const size_t row_q = (size_t)A.stride / 4;
    for (int i = 0; i < A.Nk; i++) {
        const uint32_t ai = (uint32_t)s_a[i];
        const uint4* ri = rows + (size_t)i * A.t * base1 * row_q;
#pragma unroll 8
        for (int j = 0; j < A.t; j++) {
            uint32_t d = (ai >> (32 - (j + 1) * A.basebit)) & mask;   // keyswitch.jl:63-67
            if (d) {
                uint4 v = __ldg(ri + ((size_t)j * base1 + (d - 1)) * row_q);
                acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w;   // keyswitch.jl:71-77
            }
        }
    }
    const int c0 = threadIdx.x * 4;
    int32_t* o = A.out + g * A.out_stride + A.out_offset;
    uint32_t vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        int c = c0 + e;
        if (c < A.n) o[c] = (int32_t)vals[e];
        else if (c == A.n) {
            int32_t* ob = A.out_b + g * A.out_stride + A.b_offset;
            if (A.b_mode == 0) *ob = (int32_t)((uint32_t)A.in[g * A.in_stride + A.in_b_offset] + vals[e]);
            else atomicAdd(reinterpret_cast<unsigned int*>(ob), vals[e]);
        }
    }
}

// Small batches (dependent circuits): one CTA per ciphertext would leave most SMs idle and make the key switch
// a 1 024-step chain of L2 round trips.  Here `slices` CTAs share a ciphertext: each gathers a contiguous range of
// mask positions and adds its partial sums into `out` — zero-initialised by the caller — with integer atomics
// (integer addition commutes, so the result is the same bits as keyswitch_kernel); slice 0 also adds the input b.
// b_mode 1 (MK, one launch per party): the joint b was written by the caller, every slice only subtracts.
__global__ void keyswitch_sliced_kernel(KeyswitchArgs A, unsigned long long count, int slices) {
    extern __shared__ int32_t s_a[];
    const size_t g = blockIdx.x / slices;
    const int sl = blockIdx.x % slices;
    if (g >= count) return;
    const int i0 = (int)((long long)A.Nk * sl / slices), i1 = (int)((long long)A.Nk * (sl + 1) / slices);
    const int32_t* in = A.in + g * A.in_stride + A.in_offset;
    const uint32_t prec = 1u << (32 - (1 + A.basebit * A.t));   // keyswitch.jl:58
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) s_a[i - i0] = (int32_t)((uint32_t)in[i] + prec);
    __syncthreads();
    const int base1 = (1 << A.basebit) - 1;
    const uint32_t mask = (uint32_t)base1;
    uint4 acc = make_uint4(0, 0, 0, 0);
    const uint4* rows = reinterpret_cast<const uint4*>(A.ksk) + threadIdx.x;
    const size_t row_q = (size_t)A.stride / 4;
    for (int i = i0; i < i1; i++) {
        const uint32_t ai = (uint32_t)s_a[i - i0];
        const uint4* ri = rows + (size_t)i * A.t * base1 * row_q;
#pragma unroll 8
        for (int j = 0; j < A.t; j++) {
            uint32_t d = (ai >> (32 - (j + 1) * A.basebit)) & mask;   // keyswitch.jl:63-67
            if (d) {
                uint4 v = __ldg(ri + ((size_t)j * base1 + (d - 1)) * row_q);
                acc.x -= v.x; acc.y -= v.y; acc.z -= v.z; acc.w -= v.w;   // keyswitch.jl:71-77
            }
        }
    }
    const int c0 = threadIdx.x * 4;
    unsigned int* o = reinterpret_cast<unsigned int*>(A.out + g * A.out_stride + A.out_offset);
    uint32_t vals[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int e = 0; e < 4; e++) {
        int c = c0 + e;
        if (c < A.n) atomicAdd(o + c, vals[e]);
        else if (c == A.n) {
            uint32_t v = vals[e];
            if (sl == 0 && A.b_mode == 0) v += (uint32_t)A.in[g * A.in_stride + A.in_b_offset];   // b_mode 1: b was copied by the caller
            atomicAdd(reinterpret_cast<unsigned int*>(A.out_b + g * A.out_stride + A.b_offset), v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row-wise linear combination of ciphertext batches: out = ka*x + kb*y + (0,...,0,cb)
// (gate_not gates.jl:76-79, gate_constant :91-93, the OR step of gate_mux :174, MK prologue mk_gates.jl:8-10)
__global__ void lincomb_kernel(const int32_t* __restrict__ x, const int32_t* __restrict__ y, int32_t* __restrict__ out,
                               int32_t ka, int32_t kb, int32_t cb, int width, unsigned long long total, int const_mode) {
    unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int col = (int)(idx % width);
    uint32_t v = 0;
    if (const_mode) {   // gate_constant: trivial sample, sign chosen by x[row][0] != 0
        if (col == width - 1) v = x[idx - col] != 0 ? (uint32_t)cb : 0u - (uint32_t)cb;
    } else {
        v = (uint32_t)ka * (uint32_t)x[idx];
        if (y) v += (uint32_t)kb * (uint32_t)y[idx];
        if (col == width - 1) v += (uint32_t)cb;
    }
    out[idx] = (int32_t)v;
}

}  // namespace tfhe_b200

namespace tfhe_b200 {
// Measures the shared-memory read rate with conflict-free LDS.128 (the roofline denominator of the tiled key switch,
// which is bound by the shared-memory pipe, not by HBM).
__global__ void lds_peak_kernel(double* out, int iters) {
    __shared__ double2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) { double2 v = buf[(idx + u * 64) & 1023]; acc.x += v.x; acc.y += v.y; }
        idx = (idx + 1) & 1023;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}
// Measures the FP64 FMA issue rate of the device (the roofline denominator of the transform kernels:
// MEASURED_PEAKS.json only records HBM and bf16 peaks).  8 independent FMA chains per thread.
__global__ void fp64_peak_kernel(double* out, double a, double b, int iters) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace tfhe_b200

// mk_kernels.cuh — sm_100a kernels of the multi-key (MK-TFHE) gate-bootstrapping path.
//
//   K5  mk_extern_product_step  mk_tgsw_extern_mul               mk_internals.jl:348-391
//       mk_blind_rotate_kernel  mk_gate_nand prologue + mk_bootstrap_wo_keyswitch
//                                                                 mk_gates.jl:7-12, mk_internals.jl:464-509, :88-95
//       (mk_keyswitch, mk_internals.jl:397-411, reuses keyswitch_kernel once per party)
//
// The reference inverse-transforms each of the l*(3p+1) products separately and sums integers because
// its unsplit FFT has no precision headroom (mk_internals.jl:359-366).  Here every key polynomial is
// split in 16-bit pieces (NP == 2), so the sums are taken exactly in the transform domain and only
// (p+1)*NP inverse transforms are needed per iteration; in exact arithmetic the results are identical.
//
// Accumulator registers: A (-> a'_party) and B (-> b') live in registers for the whole iteration;
// S (-> a'_i, i != party) is short-lived and kept in shared memory (each thread only ever touches its
// own 8*NP entries, so no barrier protects it).
#pragma once
#include "kernels.cuh"

namespace tfhe_b200 {

// indices of the polynomials inside one expanded MK sample (mk_internals.jl:240-250), [l][p] row-major
__host__ __device__ inline int mk_xi(int l, int p, int r, int i) { (void)l; return r * p + i; }
__host__ __device__ inline int mk_yi(int l, int p, int r, int i) { return l * p + r * p + i; }
__host__ __device__ inline int mk_c0i(int l, int p, int r) { return 2 * l * p + r; }
__host__ __device__ inline int mk_c1i(int l, int p, int r) { return 2 * l * p + l + r; }

template <int NP>
__device__ __forceinline__ void mac_spectrum(double2 (&o)[NP][8], const double2 (&a)[8], const double2* __restrict__ b, int t) {
#pragma unroll
    for (int pc = 0; pc < NP; pc++)
#pragma unroll
        for (int q = 0; q < 8; q++) cmac(o[pc][q], a[q], __ldg(b + (pc * 8 + q) * 64 + t));
}

// inverse-transform NP pieces, reassemble lo + (hi << 16), add to (or store into) coefficient poly p
template <int NP, bool ACCUM>
__device__ __forceinline__ void finish_poly(double2 (&o)[NP][8], int32_t* p, const Twiddles& w, double2* X1, double2* X2,
                                            int t, int bar_id) {
    uint32_t rl[8], rh[8];
#pragma unroll
    for (int pc = 0; pc < NP; pc++) {
        fft512_inverse(o[pc], w, X1, X2, t, bar_id);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            uint32_t vl = round_to_u32_fast<NP == 2>(o[pc][m].x), vh = round_to_u32_fast<NP == 2>(-o[pc][m].y);
            if (pc == 0) { rl[m] = vl; rh[m] = vh; }
            else { rl[m] += vl << 16; rh[m] += vh << 16; }
        }
    }
#pragma unroll
    for (int m = 0; m < 8; m++) {
        int j = t + 64 * m;
        if (ACCUM) { p[j] = (int32_t)((uint32_t)p[j] + rl[m]); p[j + 512] = (int32_t)((uint32_t)p[j + 512] + rh[m]); }
        else { p[j] = (int32_t)rl[m]; p[j + 512] = (int32_t)rh[m]; }
    }
}

// One MK external product on the accumulator acc[(p+1)][N] in shared memory (a_1..a_p, b).
// `sample` points at the spectra of bk.key[j, party]: [poly][piece][512].
template <int L, int BGBIT, int NP, bool ROTSUB, bool ACCUM>
__device__ __forceinline__ void mk_extern_product_step(int32_t* acc, int p, int party, int abar,
                                                       const double2* __restrict__ sample, const Twiddles& w,
                                                       double2* X1, double2* X2, double2* S, int t, int bar_id) {
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
    const size_t PS = (size_t)NP * kSpectrum;   // doubles2 per stored polynomial
    double2 A[NP][8], B[NP][8];
#pragma unroll
    for (int pc = 0; pc < NP; pc++)
#pragma unroll
        for (int q = 0; q < 8; q++) { A[pc][q] = make_double2(0.0, 0.0); B[pc][q] = make_double2(0.0, 0.0); }

#pragma unroll 1
    for (int q = 0; q <= p; q++) {
        const int32_t* poly = acc + q * kN;
        uint32_t tl[8], th[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            if (ROTSUB) {
                tl[m] = (uint32_t)rot_coeff(poly, j, s) - (uint32_t)poly[j] + offset;          // mk_internals.jl:468
                th[m] = (uint32_t)rot_coeff(poly, j + 512, s) - (uint32_t)poly[j + 512] + offset;
            } else {
                tl[m] = (uint32_t)poly[j] + offset;
                th[m] = (uint32_t)poly[j + 512] + offset;
            }
        }
        const bool side = q < p && q != party;   // this input also feeds a'_q through y[.,party]
        if (side) {
#pragma unroll
            for (int e = 0; e < NP * 8; e++) S[e * 64 + t] = make_double2(0.0, 0.0);
        }
#pragma unroll 1
        for (int r = 0; r < L; r++) {
            double2 a[8];
#pragma unroll
            for (int m = 0; m < 8; m++)
                a[m] = make_double2(digit_f64<BGBIT>(tl[m], r), -digit_f64<BGBIT>(th[m], r));   // :356-357
            fft512_forward(a, w, X1, X2, t, bar_id);                                                    // :368-369
            if (q < p) {
                mac_spectrum<NP>(A, a, sample + (size_t)mk_yi(L, p, r, q) * PS, t);                     // :375-376
                mac_spectrum<NP>(B, a, sample + (size_t)mk_xi(L, p, r, q) * PS, t);                     // :384-385
                if (side) {                                                                             // :379-380
                    const double2* b = sample + (size_t)mk_yi(L, p, r, party) * PS;
#pragma unroll
                    for (int e = 0; e < NP * 8; e++) {
                        double2 sv = S[e * 64 + t];
                        cmac(sv, a[e & 7], __ldg(b + e * 64 + t));
                        S[e * 64 + t] = sv;
                    }
                }
            } else {
                mac_spectrum<NP>(A, a, sample + (size_t)mk_c1i(L, p, r) * PS, t);                       // :377-378
                mac_spectrum<NP>(B, a, sample + (size_t)mk_c0i(L, p, r) * PS, t);                       // :386-387
            }
        }
        if (side) {
            double2 o[NP][8];
#pragma unroll
            for (int pc = 0; pc < NP; pc++)
#pragma unroll
                for (int e = 0; e < 8; e++) o[pc][e] = S[(pc * 8 + e) * 64 + t];
            group_sync(bar_id);   // all reads of acc[q] and of X2 (last forward) are done
            finish_poly<NP, ACCUM>(o, acc + q * kN, w, X1, X2, t, bar_id);
            group_sync(bar_id);   // X1 free again before the next forward transform
        }
    }
    group_sync(bar_id);
    finish_poly<NP, ACCUM>(A, acc + party * kN, w, X1, X2, t, bar_id);
    finish_poly<NP, ACCUM>(B, acc + p * kN, w, X1, X2, t, bar_id);
    group_sync(bar_id);
}

struct MKBlindRotateArgs {
    const double2* bk_fft;   // [p][n][L*(2p+2)][NP][512]
    const double2* E;
    const int32_t* x; const int32_t* y;   // [count][p*n+1]; lin = ka*x + kb*y + (0, cb)
    int32_t ka, kb, cb, mu;
    int32_t* out;            // [count][p*N+1]
    int n, p;
    unsigned long long count;
};

__host__ __device__ inline size_t mk_smem_bytes(int p, int n, int NP) {
    return (size_t)(kSpectrum + kX2Elems) * 16 + (size_t)NP * kSpectrum * 16 + (size_t)(p + 1) * kN * 4 + (size_t)((p * n + 3) & ~3) * 4;
}

template <int L, int BGBIT, int NP>
__global__ void __launch_bounds__(64) mk_blind_rotate_kernel(MKBlindRotateArgs M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int t = threadIdx.x, p = M.p, n = M.n;
    double2* X1 = reinterpret_cast<double2*>(smem_raw);
    double2* X2 = X1 + kSpectrum;
    double2* S = X2 + kX2Elems;
    int32_t* acc = reinterpret_cast<int32_t*>(S + NP * kSpectrum);
    int32_t* bara = acc + (p + 1) * kN;
    const size_t g = blockIdx.x;
    Twiddles w; w.load(M.E, t);
    const size_t wct = (size_t)p * n + 1;
    const int32_t* xr = M.x + g * wct;
    const int32_t* yr = M.y ? M.y + g * wct : nullptr;
    for (int i = t; i < p * n; i += 64) {                                       // mk_gates.jl:8-10, mk_internals.jl:503
        uint32_t v = (uint32_t)M.ka * (uint32_t)xr[i];
        if (yr) v += (uint32_t)M.kb * (uint32_t)yr[i];
        bara[i] = modswitch2048((int32_t)v);
    }
    uint32_t vb = (uint32_t)M.ka * (uint32_t)xr[p * n] + (uint32_t)M.cb;
    if (yr) vb += (uint32_t)M.kb * (uint32_t)yr[p * n];
    const int barb = modswitch2048((int32_t)vb);                                // :502
    const int s0 = (-barb) & 2047;
    for (int x = t; x < p * kN; x += 64) acc[x] = 0;                            // :69-76
    for (int x = t; x < kN; x += 64) {                                          // :491-492, :506
        int yy = (x - s0) & 2047;
        acc[p * kN + x] = (yy & 1024) ? (int32_t)(0u - (uint32_t)M.mu) : M.mu;
    }
    __syncthreads();
    const size_t spolys = (size_t)L * (2 * p + 2);
#pragma unroll 1
    for (int party = 0; party < p; party++)                                     // :475
#pragma unroll 1
        for (int j = 0; j < n; j++) {                                           // :476
            const int abar = bara[party * n + j];
            if (abar == 0) continue;                                            // :478
            const double2* sample = M.bk_fft + ((size_t)party * n + j) * spolys * NP * kSpectrum;
            mk_extern_product_step<L, BGBIT, NP, true, true>(acc, p, party, abar, sample, w, X1, X2, S, t, 0);
        }
    // mk_tlwe_extract_sample (mk_internals.jl:88-95)
    int32_t* o = M.out + g * ((size_t)p * kN + 1);
    for (int q = 0; q < p; q++)
        for (int x = t; x < kN; x += 64)
            o[q * kN + x] = x == 0 ? acc[q * kN] : (int32_t)(0u - (uint32_t)acc[q * kN + kN - x]);
    if (t == 0) o[p * kN] = acc[p * kN];
}

template <int L, int BGBIT, int NP>
__global__ void __launch_bounds__(64) mk_extern_product_kernel(const double2* __restrict__ bk_fft,
                                                               const double2* __restrict__ E,
                                                               const int32_t* __restrict__ acc_in,
                                                               const int32_t* __restrict__ party,
                                                               const int32_t* __restrict__ bk_index,
                                                               int32_t* __restrict__ out, int n, int p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int t = threadIdx.x;
    double2* X1 = reinterpret_cast<double2*>(smem_raw);
    double2* X2 = X1 + kSpectrum;
    double2* S = X2 + kX2Elems;
    int32_t* acc = reinterpret_cast<int32_t*>(S + NP * kSpectrum);
    const size_t g = blockIdx.x;
    Twiddles w; w.load(E, t);
    const size_t words = (size_t)(p + 1) * kN;
    for (int x = t; x < (int)words; x += 64) acc[x] = acc_in[g * words + x];
    __syncthreads();
    const size_t spolys = (size_t)L * (2 * p + 2);
    const double2* sample = bk_fft + ((size_t)party[g] * n + bk_index[g]) * spolys * NP * kSpectrum;
    mk_extern_product_step<L, BGBIT, NP, false, false>(acc, p, party[g], 0, sample, w, X1, X2, S, t, 0);
    for (int x = t; x < (int)words; x += 64) out[g * words + x] = acc[x];
}

// result.b = sample.b (mk_internals.jl:405); the per-party key switches then add their b parts (:409)
__global__ void mk_copy_b_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, long long in_stride,
                                 long long in_b, long long out_stride, long long out_b, unsigned long long count) {
    unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < count) out[g * out_stride + out_b] = in[g * in_stride + in_b];
}

}  // namespace tfhe_b200

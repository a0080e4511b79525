// blind_rotate_wide.cuh — the gate bootstrap for TLWE mask sizes k = 2 and 3.
//
// The reference's parameter constructors take `tlwe_mask_size` as a keyword (api.jl:30,55): the accumulator then has
// k+1 polynomials (tlwe.jl:36), a bootstrapping-key element is an l(k+1) x (k+1) matrix of them (tgsw.jl:28,62-69),
// decompose yields l(k+1) digit polynomials (tgsw.jl:99-117) and the extracted LWE sample has k*N mask words
// (tlwe.jl:55-59, 30).  No test or example of the reference uses k > 1, so this path is built for parity, not tuned:
// one 64-thread group per gate, four gates per CTA, the key read straight from L2 (software-pipelined), and the (k+1)*NP
// output spectra — 96 or 128 registers' worth too many for the register file next to a transform — accumulated in
// TENSOR MEMORY (tmem.cuh: every thread owns a TMEM row; read-modify-write on the LDTM/STTM datapath, off the
// shared-memory pipe; 31-35 KB of shared memory per gate instead of 78-98 KB, so four gates share an SM instead of two:
// 21.0 k -> 34.2 k gates/s at k = 2).  The stand-alone external product keeps them in shared memory.
// Arithmetic, summation order and rounding are those of extern_product_step (kernels.cuh): the magnitudes grow by
// (k+1)/2 (<= 2^37 for k = 3 with two 16-bit key pieces), far inside the 2^41 the rounding trick is proven for.
#pragma once
#include "kernels.cuh"
#include "blind_rotate.cuh"

namespace tfhe_b200 {

constexpr int kWideGates = 4;   // gates (64-thread groups) per CTA: 8 warps, two per TMEM lane quarter, 256 columns each

// per gate: X1 + X2 exchange buffers, [(k+1)*NP output spectra: stand-alone external product only], the accumulator, bara
__host__ __device__ constexpr size_t br_wide_group_bytes(int NP, int kp1, int n_pad, bool out_in_smem) {
    return (size_t)(kSpectrum + kX2Elems) * 16 + (out_in_smem ? (size_t)kp1 * NP * kSpectrum * 16 : 0) + (size_t)kp1 * kN * 4 + (size_t)n_pad * 4;
}
__host__ __device__ constexpr size_t br_wide_smem_bytes(int NP, int kp1, int n_pad) {
    return (size_t)kWideGates * br_wide_group_bytes(NP, kp1, n_pad, false);
}

// where the output spectra of an external product live: spectrum o, this thread's 8 points
struct WideOutSmem {
    double2* O; int t;
    __device__ __forceinline__ void load(int o, double2 (&v)[8]) const {
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = O[o * kSpectrum + q * 64 + t];
    }
    __device__ __forceinline__ void store(int o, const double2 (&v)[8]) const {
#pragma unroll
        for (int q = 0; q < 8; q++) O[o * kSpectrum + q * 64 + t] = v[q];
    }
    __device__ __forceinline__ void stores_landed() const {}
};
struct WideOutTmem {
    uint32_t tm;   // TMEM address of this thread's row, first column of its warp's range; spectrum o = columns [32 o, 32 o + 32)
    __device__ __forceinline__ void load(int o, double2 (&v)[8]) const { tmem_load_spectrum(tm + (uint32_t)(o * 32), v); }
    __device__ __forceinline__ void store(int o, const double2 (&v)[8]) const { tmem_store_spectrum(tm + (uint32_t)(o * 32), v); }
    __device__ __forceinline__ void stores_landed() const { tmem_wait_st(); }
};

// One external product on an accumulator of kp1 polynomials in shared memory (the k-generic form of extern_product_step):
//   temp_c = ROTSUB ? X^abar * acc_c - acc_c : acc_c
//   acc_c' = (ACCUM ? acc_c' : 0) + sum_{c, r} digit_r(temp_c) (*) BK[r][c][c']
// bk_row: this key element's spectra [r][c][c'][piece][512].  O: kp1*NP spectra of scratch.
template <int L, int BGBIT, int NP, bool ROTSUB, bool ACCUM, class OUT>
__device__ __forceinline__ void extern_product_step_wide(int32_t* acc, int kp1, int abar, const double2* __restrict__ bk_row,
                                                         const Twiddles& w, double2* X1, double2* X2, const OUT& O, int t, int bar_id) {
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
    const int nout = kp1 * NP;
#pragma unroll 1
    for (int c = 0; c < kp1; c++) {
        const int32_t* p = acc + c * kN;
        uint32_t tl[8], th[8];   // temp_c at j and j+512, offset already added
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int j = t + 64 * m;
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
            // the key spectra come straight from L2 (~800 cycles): the 8 points of output 0 are requested before the
            // transform and those of output o+1 before the multiply-accumulate of output o, so a load is never waited for
            const double2* b = bk_row + (size_t)(r * kp1 + c) * nout * kSpectrum + t;
            double2 kb[8];
#pragma unroll
            for (int q = 0; q < 8; q++) kb[q] = __ldg(b + q * 64);
            fft512_forward(a, w, X1, X2, t, bar_id);
            const bool first = c == 0 && r == 0;
            if (!first) O.stores_landed();   // this thread's accumulator stores of the previous digit polynomial
#pragma unroll 1
            for (int o = 0; o < nout; o++) {   // o = c' * NP + piece
                double2 kn[8], v[8];
                const double2* bn = b + (size_t)(o + 1 < nout ? o + 1 : o) * kSpectrum;
#pragma unroll
                for (int q = 0; q < 8; q++) kn[q] = __ldg(bn + q * 64);
                if (first) {
#pragma unroll
                    for (int q = 0; q < 8; q++) v[q] = make_double2(0.0, 0.0);
                } else {
                    O.load(o, v);
                }
#pragma unroll
                for (int q = 0; q < 8; q++) cmac(v[q], a[q], kb[q]);
                O.store(o, v);
#pragma unroll
                for (int q = 0; q < 8; q++) kb[q] = kn[q];
            }
        }
    }
    O.stores_landed();
    // every thread has finished reading acc and X2 (last forward) before anyone overwrites them
    group_sync(bar_id);
#pragma unroll 1
    for (int c2 = 0; c2 < kp1; c2++) {
        uint32_t rl[8], rh[8];
#pragma unroll
        for (int pc = 0; pc < NP; pc++) {
            double2 oc[8];
            O.load(c2 * NP + pc, oc);
            fft512_inverse(oc, w, X1, X2, t, bar_id);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                uint32_t vl = round_to_u32_fast<NP == 2>(oc[m].x), vh = round_to_u32_fast<NP == 2>(-oc[m].y);
                if (pc == 0) { rl[m] = vl; rh[m] = vh; }
                else { rl[m] += vl << 16; rh[m] += vh << 16; }
            }
        }
        int32_t* p = acc + c2 * kN;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int j = t + 64 * m;
            if (ACCUM) { p[j] = (int32_t)((uint32_t)p[j] + rl[m]); p[j + 512] = (int32_t)((uint32_t)p[j + 512] + rh[m]); }
            else { p[j] = (int32_t)rl[m]; p[j + 512] = (int32_t)rh[m]; }
        }
    }
    group_sync(bar_id);
}

struct WideGroup {
    double2 *X1, *X2, *O; int32_t *acc, *bara;
    __device__ __forceinline__ WideGroup(unsigned char* base, int NP, int kp1, bool out_in_smem) {
        X1 = reinterpret_cast<double2*>(base);
        X2 = X1 + kSpectrum;
        O = X2 + kX2Elems;
        acc = reinterpret_cast<int32_t*>(O + (out_in_smem ? (size_t)kp1 * NP * kSpectrum : 0));
        bara = acc + kp1 * kN;
    }
};

// K2 for k > 1 as a stand-alone batch kernel (parity tests of tgsw_extern_mul): one group per product
template <int L, int BGBIT, int NP>
__global__ void __launch_bounds__(64) extern_product_wide_kernel(const double2* __restrict__ bk_fft, const double2* __restrict__ E,
                                                                 const int32_t* __restrict__ acc_in, const int32_t* __restrict__ bk_index,
                                                                 int32_t* __restrict__ out, int kp1) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WideGroup G(smem_raw, NP, kp1, true);
    const int t = threadIdx.x;
    Twiddles w; w.load(E, t);
    const size_t g = blockIdx.x, words = (size_t)kp1 * kN;
    for (int x = t; x < (int)words; x += 64) G.acc[x] = acc_in[g * words + x];
    __syncthreads();
    const size_t row = (size_t)L * kp1 * kp1 * NP * kSpectrum;
    extern_product_step_wide<L, BGBIT, NP, false, false>(G.acc, kp1, 0, bk_fft + (size_t)bk_index[g] * row, w, G.X1, G.X2, WideOutSmem{G.O, t}, t, 0);
    for (int x = t; x < (int)words; x += 64) out[g * words + x] = G.acc[x];
}

// K3 for k > 1.  MODE 0: gate prologue + modulus switch + blind rotation + sample extraction (out [count][k*N + 1]);
// MODE 1: blind_rotate of given accumulators (acc_in / out [count][k+1][N], bara_in [count][n]).
template <int L, int BGBIT, int NP, int MODE>
__global__ void __launch_bounds__(64 * kWideGates) blind_rotate_wide_kernel(BlindRotateArgs A, int kp1) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem_base;
    // all 512 TMEM columns (one CTA per SM): warps that share a lane quarter (warp % 4) take 256 columns each, enough
    // for the (k+1)*NP <= 8 output spectra of 32 columns
    if ((threadIdx.x >> 5) == 0) tmem_alloc<512>(&s_tmem_base);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6, warp = threadIdx.x >> 5;
    const int bar_id = grp + 1;
    const WideOutTmem O{s_tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 256)};
    WideGroup G(smem_raw + (size_t)grp * br_wide_group_bytes(NP, kp1, A.n_pad, false), NP, kp1, false);
    // gate -> CTA map as in K3 (BlindRotateArgs::split / tail): the last wave of CTAs spreads its gates over all SMs
    const int cta_gates = blockIdx.x < A.split ? kWideGates : A.tail;
    const unsigned long long g_first = blockIdx.x < A.split ? (unsigned long long)blockIdx.x * kWideGates
                                                            : (unsigned long long)A.split * kWideGates + (unsigned long long)(blockIdx.x - A.split) * A.tail;
    const unsigned long long g = g_first + grp;
    const bool valid = grp < cta_gates && g < A.count;   // a group without a gate only waits for the others at the end (TMEM is freed by warp 0)
    Twiddles w; w.load(A.E, t);
    const int k = kp1 - 1;
    int32_t* acc = G.acc;
    int32_t* bara = G.bara;

    if (!valid) {
    } else if (MODE == 0) {
        // gate prologue (gates.jl) + modulus switch (bootstrap.jl:74-75)
        const bool second = A.half != 0 && g >= A.half;
        const unsigned long long gi = second ? g - A.half : g;
        const int32_t* xr = (second ? A.x2 : A.x) + gi * (A.n + 1);
        const int32_t* yb = second ? A.y2 : A.y;
        const int32_t* yr = yb ? yb + gi * (A.n + 1) : nullptr;
        const int32_t ka = second ? A.ka2 : A.ka, kb = second ? A.kb2 : A.kb, cb = second ? A.cb2 : A.cb;
        for (int i = t; i < A.n; i += 64) {
            uint32_t v = (uint32_t)ka * (uint32_t)xr[i];
            if (yr) v += (uint32_t)kb * (uint32_t)yr[i];
            bara[i] = modswitch2048((int32_t)v);
        }
        uint32_t vb = (uint32_t)ka * (uint32_t)xr[A.n] + (uint32_t)cb;
        if (yr) vb += (uint32_t)kb * (uint32_t)yr[A.n];
        const int barb = modswitch2048((int32_t)vb);
        // acc = (0, ..., 0, X^{-barb} * (mu, ..., mu))   (bootstrap.jl:54-56,78)
        const int s = (-barb) & 2047;
        for (int x = t; x < k * kN; x += 64) acc[x] = 0;
        for (int x = t; x < kN; x += 64) {
            const int yy = (x - s) & 2047;
            acc[k * kN + x] = (yy & 1024) ? (int32_t)(0u - (uint32_t)A.mu) : A.mu;
        }
    } else {
        const int32_t* ain = A.acc_in + g * ((size_t)kp1 * kN);
        for (int x = t; x < kp1 * kN; x += 64) acc[x] = ain[x];
        for (int i = t; i < A.n_iter; i += 64) bara[i] = A.bara_in[g * A.n + i];
    }
    group_sync(bar_id);

    const size_t row = (size_t)L * kp1 * kp1 * NP * kSpectrum;
    const int n_walk = __all_sync(0xffffffffu, valid) ? A.n_iter : 0;   // warp-uniform for the compiler (see blind_rotate.cuh)
#pragma unroll 1
    for (int i = 0; i < n_walk; i++)   // bootstrap.jl:19-23
        extern_product_step_wide<L, BGBIT, NP, true, true>(acc, kp1, bara[i], A.bk_fft + (size_t)i * row, w, G.X1, G.X2, O, t, bar_id);

    if (!valid) {
    } else if (MODE == 0) {
        // tlwe_extract_sample (tlwe.jl:55-59): a_j = (p_j0, -p_j,N-1, ..., -p_j1) for every mask polynomial, b = acc_k[0]
        int32_t* o = A.out + g * ((size_t)k * kN + 1);
        for (int x = t; x < k * kN; x += 64) {
            const int j = x >> 10, xx = x & (kN - 1);
            o[x] = xx == 0 ? acc[j * kN] : (int32_t)(0u - (uint32_t)acc[j * kN + kN - xx]);
        }
        if (t == 0) o[k * kN] = acc[k * kN];
    } else {
        int32_t* o = A.out + g * ((size_t)kp1 * kN);
        for (int x = t; x < kp1 * kN; x += 64) o[x] = acc[x];
    }
    tmem_fence_before_sync();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc<512>(s_tmem_base);
}

}  // namespace tfhe_b200

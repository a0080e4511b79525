// mk_blind_rotate_lowlat.cuh — K5L: MK-TFHE blind rotation of ONE gate per CTA for two parties
// (examples/multikey.jl works on a handful of bits, so what counts there is the latency of one mk_bootstrap).
//
// Same idea as K3L (blind_rotate_lowlat.cuh), for mk_tgsw_extern_mul (mk_internals.jl:348-391) with p = 2:
// the (p+1)*l = 12 digit polynomials of an iteration are spread over 6 groups of 64 threads,
//
//   phase 1  group g: polynomial q = g / 2 of the accumulator (a_1, a_2, b), digits r = 2*(g % 2) + {0, 1}:
//            rotate/subtract (:468), decompose (:356-357), forward transform, publish F[q*l + r]   -> barrier A
//   phase 2  group g = (kind, piece): kind 0: A -> a'_party   = sum_d F[d] * (y[r,q] | c1[r])      (:375-378)
//                                     kind 1: B -> b'         = sum_d F[d] * (x[r,q] | c0[r])      (:384-387)
//                                     kind 2: S -> a'_other   = sum_r F[other*l + r] * y[r,party]  (:379-380)
//            in the summation order of mk_kernels.cuh (q outer, r inner), key spectra read straight from L2
//                                                                                                   -> barrier B
//            inverse transform, round, high piece << 16, integer atomic add into the accumulator    -> barrier C
//
// With one 32-bit piece (NP == 1) only groups 0, 2, 4 work in phase 2.  Iterations whose rotation is zero are
// skipped as in the reference (:478); the CTA holds one gate, so the branch is uniform.
#pragma once
#include "mk_kernels.cuh"

namespace tfhe_b200 {

constexpr int kMkLowlatGroups = 6;
__host__ __device__ inline size_t mk_lowlat_smem_bytes(int L, int n) {
    const int p = 2;
    return (size_t)(p + 1) * L * kSpectrum * 16                      // published spectra F[(p+1)*l]
           + (size_t)kMkLowlatGroups * (kSpectrum + kX2Elems) * 16   // X1, X2 per group
           + (size_t)(p + 1) * kN * 4 + (((size_t)p * n * 2 + 15) & ~(size_t)15);
}

template <int L, int BGBIT, int NP>
__global__ void __launch_bounds__(64 * kMkLowlatGroups, 1) mk_blind_rotate_lowlat_kernel(MKBlindRotateArgs M) {
    static_assert(L % 2 == 0, "two digits per group");
    constexpr int p = 2, NG = kMkLowlatGroups, ND = (p + 1) * L;
    static_assert(ND == 2 * NG, "(p+1)*l digit polynomials over 6 groups, two each");
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* F = reinterpret_cast<double2*>(smem_raw);
    double2* xbuf = F + (size_t)ND * kSpectrum;
    int32_t* acc = reinterpret_cast<int32_t*>(xbuf + (size_t)NG * (kSpectrum + kX2Elems));
    int16_t* bara = reinterpret_cast<int16_t*>(acc + (p + 1) * kN);

    const int n = M.n;
    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int bar_id = grp + 1;
    double2* X1 = xbuf + (size_t)grp * (kSpectrum + kX2Elems);
    double2* X2 = X1 + kSpectrum;
    const size_t g = blockIdx.x;
    Twiddles w; w.load(M.E, t);

    {   // mk_gates.jl:8-10 prologue + modulus switch (mk_internals.jl:502-503) + test vector (:491-492, :506)
        const size_t wct = (size_t)p * n + 1;
        const int32_t* xr = M.x + g * wct;
        const int32_t* yr = M.y ? M.y + g * wct : nullptr;
        for (int i = threadIdx.x; i < p * n; i += blockDim.x) {
            uint32_t v = (uint32_t)M.ka * (uint32_t)xr[i];
            if (yr) v += (uint32_t)M.kb * (uint32_t)yr[i];
            bara[i] = (int16_t)modswitch2048((int32_t)v);
        }
        uint32_t vb = (uint32_t)M.ka * (uint32_t)xr[p * n] + (uint32_t)M.cb;
        if (yr) vb += (uint32_t)M.kb * (uint32_t)yr[p * n];
        const int barb = modswitch2048((int32_t)vb);
        const int s0 = (-barb) & 2047;
        for (int x = threadIdx.x; x < p * kN; x += blockDim.x) acc[x] = 0;
        for (int x = threadIdx.x; x < kN; x += blockDim.x) {
            int yy = (x - s0) & 2047;
            acc[p * kN + x] = (yy & 1024) ? (int32_t)(0u - (uint32_t)M.mu) : M.mu;
        }
    }
    __syncthreads();

    const int q1 = grp >> 1, r0 = 2 * (grp & 1);          // phase-1 role
    const int kind = grp / NP, pc = grp % NP;             // phase-2 role (NP == 2: all groups; NP == 1: kind = grp, 3 groups)
    const bool worker = NP == 2 || grp < 3;
    const int32_t* poly = acc + q1 * kN;
    const size_t PS = (size_t)NP * kSpectrum, spolys = (size_t)L * (2 * p + 2);

#pragma unroll 1
    for (int party = 0; party < p; party++)                                     // mk_internals.jl:475
#pragma unroll 1
        for (int j = 0; j < n; j++) {                                           // :476
            const int abar = bara[party * n + j];
            if (abar == 0) continue;                                            // :478 (uniform: one gate per CTA)
            const int s = abar & 2047;
            const int other = 1 - party;
            const double2* sample = M.bk_fft + ((size_t)party * n + j) * spolys * PS + (size_t)pc * kSpectrum + t;
            // step d of this group's sum: which published spectrum, which key polynomial (order of mk_kernels.cuh)
            const int nsteps = kind < 2 ? ND : L;
            auto f_index = [&](int d) { return kind < 2 ? d : other * L + d; };
            auto key_poly = [&](int d) {
                const int q = d / L, r = d % L;
                return kind == 0 ? (q < p ? mk_yi(L, p, r, q) : mk_c1i(L, p, r))
                     : kind == 1 ? (q < p ? mk_xi(L, p, r, q) : mk_c0i(L, p, r))
                                 : mk_yi(L, p, d, party);
            };
            if (worker && t == 0) {   // the expanded key does not fit L2: pull this iteration's spectra in (one bulk
#pragma unroll                 // prefetch of 8 KB per step) while phase 1 computes
                for (int d = 0; d < ND; d++)
                    if (d < nsteps)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                                     ::"l"(sample + (size_t)key_poly(d) * PS), "n"(kSpectrum * 16) : "memory");
            }
            {
                uint32_t tl[8], th[8];
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int c = t + 64 * m;
                    tl[m] = (uint32_t)rot_coeff(poly, c, s) - (uint32_t)poly[c] + offset;               // :468
                    th[m] = (uint32_t)rot_coeff(poly, c + 512, s) - (uint32_t)poly[c + 512] + offset;
                }
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int r = r0 + rr;
                    double2 a[8];
#pragma unroll
                    for (int m = 0; m < 8; m++)
                        a[m] = make_double2(digit_f64<BGBIT>(tl[m], r), -digit_f64<BGBIT>(th[m], r));  // :356-357
                    fft512_forward(a, w, X1, X2, t, bar_id);                                            // :368-369
                    double2* Fd = F + (size_t)(q1 * L + r) * kSpectrum + t;
#pragma unroll
                    for (int e = 0; e < 8; e++) Fd[e * 64] = a[e];
                }
            }
            __syncthreads();   // A: all 12 spectra published, all reads of the accumulator done
            double2 o[8];
            if (worker) {
#pragma unroll
                for (int e = 0; e < 8; e++) o[e] = make_double2(0.0, 0.0);
                // explicit double buffering: the key spectrum of step d+1 is in flight while step d multiplies
                double2 kv[8];
#pragma unroll
                for (int e = 0; e < 8; e++) kv[e] = __ldg(sample + (size_t)key_poly(0) * PS + e * 64);
#pragma unroll
                for (int d = 0; d < ND; d++) {
                    if (d < nsteps) {
                        double2 kn[8];
                        if (d + 1 < nsteps) {
#pragma unroll
                            for (int e = 0; e < 8; e++) kn[e] = __ldg(sample + (size_t)key_poly(d + 1) * PS + e * 64);
                        }
                        const double2* Fd = F + (size_t)f_index(d) * kSpectrum + t;
#pragma unroll
                        for (int e = 0; e < 8; e++) cmac(o[e], Fd[e * 64], kv[e]);
                        if (d + 1 < nsteps) {
#pragma unroll
                            for (int e = 0; e < 8; e++) kv[e] = kn[e];
                        }
                    }
                }
            }
            __syncthreads();   // B: published spectra consumed
            if (worker) {
                fft512_inverse(o, w, X1, X2, t, bar_id);
                int32_t* pa = acc + (kind == 0 ? party : kind == 1 ? p : other) * kN;
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    uint32_t vl = round_to_u32_fast<NP == 2>(o[m].x), vh = round_to_u32_fast<NP == 2>(-o[m].y);
                    if (pc == 1) { vl <<= 16; vh <<= 16; }
                    const int c = t + 64 * m;
                    atomicAdd(reinterpret_cast<unsigned int*>(pa + c), vl);
                    atomicAdd(reinterpret_cast<unsigned int*>(pa + c + 512), vh);
                }
            }
            __syncthreads();   // C: accumulator updated
        }

    // mk_tlwe_extract_sample (mk_internals.jl:88-95)
    int32_t* out = M.out + g * ((size_t)p * kN + 1);
    for (int q = 0; q < p; q++)
        for (int x = threadIdx.x; x < kN; x += blockDim.x)
            out[q * kN + x] = x == 0 ? acc[q * kN] : (int32_t)(0u - (uint32_t)acc[q * kN + kN - x]);
    if (threadIdx.x == 0) out[p * kN] = acc[p * kN];
}

}  // namespace tfhe_b200

// blind_rotate_lowlat.cuh — K3L: blind rotation of ONE gate per CTA, parallel across the digit polynomials.
//
// K3 (blind_rotate.cuh) maximises throughput: one 64-thread group runs all 4 forward and 2*NP inverse transforms
// of an iteration one after the other.  A dependent circuit (examples/tutorial.jl, a ripple-carry adder) has
// one or two gates per level, so what counts there is the latency of a single bootstrap: 500 strictly sequential
// iterations.  This kernel spreads one iteration over (k+1)*l = 4 groups:
//
//   phase 1  group q = (c, r): rotate/subtract acc[c], extract digit r (tgsw.jl:104-116), forward transform,
//            publish the spectrum F_q in its exchange buffer                       -> CTA barrier A
//   phase 2  group o = (c', piece), o < 2*NP: O = sum_q F_q * BK[i][r][c][c'][piece] (tgsw.jl:128), the 4 key
//            spectra having been prefetched into shared memory by TMA during the previous iteration
//                                                                                  -> CTA barrier B
//            inverse transform, round (polynomials.jl:115-116), shift the high piece by 16 and add into acc[c']
//            with shared-memory integer atomics (integer addition commutes: deterministic) -> CTA barrier C
//
// so the critical path of an iteration is one forward transform, 4 spectrum MACs and one inverse transform
// instead of 4 + 16 + 4.  Results are bit-identical to K3 (same transforms, the sums over q are taken in the
// same order).  Used for batches of up to three gates per SM (cabi.cu, launch_br_np).
#pragma once
#include "blind_rotate.cuh"

namespace tfhe_b200 {

// Key slots per output group.  All 2l key spectra of an output are prefetched during the previous iteration when that
// fits (80-bit set, either mode; 128-bit set with one piece).  The 128-bit set with two pieces would need 4 x 6 x 8 KB
// = 192 KB of key slots beside 102 KB of exchange buffers: there every output group walks its 6 spectra through a ring
// of 3 slots, refilling a slot as soon as the whole group has consumed it (one extra 64-thread barrier per refill).
template <int L, int NP> __host__ __device__ constexpr int br_lowlat_key_slots() {
    return ((2 * NP) * (2 * L) * 8 + (2 * L) * 17 + 12 <= 227) ? 2 * L : 3;
}
template <int L, int NP> __host__ __device__ constexpr size_t br_lowlat_smem_bytes(int n_pad) {
    return (size_t)(2 * NP) * br_lowlat_key_slots<L, NP>() * kSpectrum * 16   // key slots [output][slot]
           + 256                                         // mbarriers [output][slot]
           + (size_t)(2 * L) * (kSpectrum + kX2Elems) * 16   // X1, X2 per group
           + 2 * kN * 4 + (size_t)n_pad * 4;             // accumulator, modulus-switched mask
}

// gate prologue (gates.jl) + modulus switch (bootstrap.jl:74-75) + initial accumulator, all threads of the CTA
__device__ __forceinline__ void lowlat_prologue(const BlindRotateArgs& A, unsigned long long g, int32_t* acc, int32_t* bara) {
    const bool second = A.half != 0 && g >= A.half;   // second bootstrap of gate_mux
    const unsigned long long gi = second ? g - A.half : g;
    const int32_t* xr = (second ? A.x2 : A.x) + gi * (A.n + 1);
    const int32_t* yb = second ? A.y2 : A.y;
    const int32_t* yr = yb ? yb + gi * (A.n + 1) : nullptr;
    const int32_t ka = second ? A.ka2 : A.ka, kb = second ? A.kb2 : A.kb, cb = second ? A.cb2 : A.cb;
    for (int i = threadIdx.x; i < A.n; i += blockDim.x) {
        uint32_t v = (uint32_t)ka * (uint32_t)xr[i];
        if (yr) v += (uint32_t)kb * (uint32_t)yr[i];
        bara[i] = modswitch2048((int32_t)v);
    }
    uint32_t vb = (uint32_t)ka * (uint32_t)xr[A.n] + (uint32_t)cb;
    if (yr) vb += (uint32_t)kb * (uint32_t)yr[A.n];
    const int barb = modswitch2048((int32_t)vb);
    const int s0 = (-barb) & 2047;   // acc = (0, X^{-barb} * (mu, ..., mu))   (bootstrap.jl:54-56,78)
    for (int x = threadIdx.x; x < kN; x += blockDim.x) {
        acc[x] = 0;
        int yy = (x - s0) & 2047;
        acc[kN + x] = (yy & 1024) ? (int32_t)(0u - (uint32_t)A.mu) : A.mu;
    }
}

template <int L, int BGBIT, int NP>
__global__ void __launch_bounds__(64 * 2 * L, 1) blind_rotate_lowlat_kernel(BlindRotateArgs A) {
    constexpr int NG = 2 * L;    // groups = digit polynomials (c, r)
    constexpr int NO = 2 * NP;   // output spectra (c', piece)
    static_assert(NO <= NG, "needs at least as many groups as output spectra");
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int KS = br_lowlat_key_slots<L, NP>();   // key slots per output group (== NG: everything prefetched)
    static_assert(NO * KS * 8 <= 256, "mbarrier area");
    double2* keys = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NO * KS * kSpectrum * 16);   // [output][slot]
    double2* xbuf = reinterpret_cast<double2*>(smem_raw + (size_t)NO * KS * kSpectrum * 16 + 256);
    int32_t* acc = reinterpret_cast<int32_t*>(xbuf + (size_t)NG * (kSpectrum + kX2Elems));
    int32_t* bara = acc + 2 * kN;

    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int bar_id = grp + 1;
    double2* X1 = xbuf + (size_t)grp * (kSpectrum + kX2Elems);
    double2* X2 = X1 + kSpectrum;
    const unsigned long long g = blockIdx.x;
    // four groups (256 threads) leave 255 registers per thread: room for the full twiddle set (no twiddle is re-derived
    // inside the two transforms of the critical path); six groups (l = 3, 384 threads, 168 registers) keep the compact set
    typename std::conditional<L == 2, TwiddlesFull, Twiddles>::type w; w.load(A.E, t);

    // output group o = (c2, pc) reads, for q = (c, r), the spectrum BK[i][r][c][c2][pc]
    const int c2 = grp / NP, pc = grp % NP;
    // spectrum number seq = i*NG + q of this output group goes to slot seq % KS; its barrier completes for the
    // (seq / KS)-th time when the copy lands
    const int total = A.n_iter * NG;
    auto issue_key = [&](int seq) {
        const int i = seq / NG, q = seq % NG, slot = seq % KS;
        const int c = q / L, r = q % L;
        const double2* src = A.bk_fft + ((((size_t)i * L + r) * 2 + c) * 2 * NP + (size_t)c2 * NP + pc) * kSpectrum;
        mbar_arrive_expect_tx(full + grp * KS + slot, (uint32_t)(kSpectrum * 16));
        bulk_copy_g2s(keys + ((size_t)grp * KS + slot) * kSpectrum, src, kSpectrum * 16, full + grp * KS + slot);
    };
    if (threadIdx.x < NO * KS) mbar_init(full + threadIdx.x, 1);
    if (threadIdx.x == 0) mbar_fence_init();
    __syncthreads();
    if (grp < NO && t == 0)
        for (int seq = 0; seq < KS && seq < total; seq++) issue_key(seq);

    lowlat_prologue(A, g, acc, bara);
    __syncthreads();

    const int c1 = grp / L, r1 = grp % L;   // phase-1 role
    const int32_t* p = acc + c1 * kN;
#pragma unroll 1
    for (int i = 0; i < A.n_iter; i++) {   // bootstrap.jl:19-23; a zero rotation is executed (exact no-op)
        const int s = bara[i] & 2047;
        {
            double2 a[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                const int j = t + 64 * m;
                const uint32_t tl = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;              // bootstrap.jl:21
                const uint32_t th = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
                a[m] = make_double2(digit_f64<BGBIT>(tl, r1), -digit_f64<BGBIT>(th, r1));                // tgsw.jl:104-116
            }
            fft512_forward(a, w, X1, X2, t, bar_id);
#pragma unroll
            for (int e = 0; e < 8; e++) X1[e * 64 + t] = a[e];   // X1 is free: every thread of the group passed the 2nd barrier
        }
        __syncthreads();   // A: all spectra published, all reads of acc done
        double2 o[8];
        if (grp < NO) {
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < NG; q++) {   // same summation order as K3: c outer, r inner
                const int seq = i * NG + q, slot = KS == NG ? q : seq % KS;
                mbar_wait(full + grp * KS + slot, (uint32_t)(KS == NG ? i : seq / KS) & 1u);
                const double2* F = xbuf + (size_t)q * (kSpectrum + kX2Elems) + t;
                const double2* K = keys + ((size_t)grp * KS + slot) * kSpectrum + t;
#pragma unroll
                for (int e = 0; e < 8; e++) cmac(o[e], F[e * 64], K[e * 64]);                            // tgsw.jl:128
                if (KS < NG) {   // ring of slots: refill this one as soon as the whole group has consumed it
                    group_sync(bar_id);
                    if (t == 0 && seq + KS < total) issue_key(seq + KS);
                }
            }
        }
        __syncthreads();   // B: spectra and key slots consumed
        if (grp < NO) {
            if (KS == NG && t == 0 && i + 1 < A.n_iter)   // lands during the inverse + next forward transform
                for (int q = 0; q < NG; q++) issue_key((i + 1) * NG + q);
            fft512_inverse(o, w, X1, X2, t, bar_id);
            int32_t* pa = acc + c2 * kN;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                uint32_t vl = round_to_u32_fast<NP == 2>(o[m].x), vh = round_to_u32_fast<NP == 2>(-o[m].y);   // polynomials.jl:115-116
                if (pc == 1) { vl <<= 16; vh <<= 16; }
                const int j = t + 64 * m;
                if (NP == 1) {
                    pa[j] = (int32_t)((uint32_t)pa[j] + vl);                                              // bootstrap.jl:22
                    pa[j + 512] = (int32_t)((uint32_t)pa[j + 512] + vh);
                } else {
                    atomicAdd(reinterpret_cast<unsigned int*>(pa + j), vl);
                    atomicAdd(reinterpret_cast<unsigned int*>(pa + j + 512), vh);
                }
            }
        }
        __syncthreads();   // C: accumulator updated
    }

    // tlwe_extract_sample (tlwe.jl:55-59): a = (p_0, -p_{N-1}, ..., -p_1), b = acc_b[0]
    int32_t* out = A.out + g * (kN + 1);
    for (int x = threadIdx.x; x < kN; x += blockDim.x) out[x] = x == 0 ? acc[0] : (int32_t)(0u - (uint32_t)acc[kN - x]);
    if (threadIdx.x == 0) out[kN] = acc[kN];
}

}  // namespace tfhe_b200

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

// ---- K3L2: the l = 2 latency kernel with the spectrum MACs taken off the shared-memory pipe --------------------------
//
// In the kernel above the MAC phase of an iteration is bound by shared-memory bandwidth: 4 output groups x (4 F + 4 key
// spectra) x 8 KB = 256 KB at 128 B/clk = 2 k of the 6.8 k cycles of an iteration (clock64 probe).  Here
//   * a group's OWN spectrum F_q stays in its registers through barrier A,
//   * the spectrum of the group that shares its tensor-memory lanes (warps w and w+4 address the same 32 lanes, so group
//     g pairs with g ^ 2; thread t of one is thread t of the other) comes through TMEM: tcgen05.st before the barrier,
//     tcgen05.ld after it, on a datapath of its own,
//   * KT = 1: the key spectra do not pass through shared memory at all.  Every thread reads exactly the key elements it
//     will multiply (K[e*64 + t]) with coalesced 16-byte read-only loads, one spectrum (8 loads) at a time, issued at the
//     start of a phase of the PREVIOUS part of the iteration (inverse transform / accumulator update / forward
//     transform) and parked in the thread's own TMEM row at the end of that phase; the fourth spectrum is consumed
//     straight from the registers it was loaded into.  KT = 0 keeps the TMA ring of the kernel above.
// Left for shared memory: two F spectra per output group (64 KB per iteration instead of 256 KB).  Same transforms, same
// summation order over q, same rounding: bit-identical to K3 and K3L.
template <int NP, int KT> __host__ __device__ constexpr size_t br_lowlat2_smem_bytes(int n_pad) {
    return (KT ? (size_t)0 : (size_t)(2 * NP) * 4 * kSpectrum * 16 + 256)   // KT = 0: key slots [output][q] + mbarriers
           + (size_t)4 * (kSpectrum + kX2Elems) * 16                       // X1, X2 per group
           + 2 * kN * 4 + (size_t)n_pad * 4;                                // accumulator, modulus-switched mask
}

template <int BGBIT, int NP, int KT, int PROBE = 0>
__global__ void __launch_bounds__(256, 1) blind_rotate_lowlat2_kernel(BlindRotateArgs A) {
    constexpr int L = 2, NG = 4, NO = 2 * NP;
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    constexpr int kTmemCols = KT ? 256 : 64;   // [0,64): F of the lower / upper group of a lane pair; [64,256): 2 x 3 key slots
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem_base;
    constexpr size_t key_bytes = KT ? 0 : (size_t)NO * NG * kSpectrum * 16 + 256;
    double2* keys = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (KT ? 0 : (size_t)NO * NG * kSpectrum * 16));   // [output][q]
    double2* xbuf = reinterpret_cast<double2*>(smem_raw + key_bytes);
    int32_t* acc = reinterpret_cast<int32_t*>(xbuf + (size_t)NG * (kSpectrum + kX2Elems));
    int32_t* bara = acc + 2 * kN;

    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6, warp = threadIdx.x >> 5;
    const int bar_id = grp + 1;
    double2* X1 = xbuf + (size_t)grp * (kSpectrum + kX2Elems);
    double2* X2 = X1 + kSpectrum;
    const unsigned long long g = blockIdx.x;
    Twiddles w; w.load(A.E, t);

    const bool outg = grp < NO;                         // this group also owns an output spectrum
    const int c2 = grp / NP, pc = grp % NP;             // ... namely (c2, pc): reads BK[i][r][c][c2][pc] for q = (c, r)
    auto key_ptr = [&](int i, int q) {
        const int c = q / L, r = q % L;
        return A.bk_fft + ((((size_t)i * L + r) * 2 + c) * 2 * NP + (size_t)c2 * NP + pc) * kSpectrum + t;
    };
    auto issue_keys = [&](int i) {   // KT = 0: the four spectra of iteration i into this output group's slots
        for (int q = 0; q < NG; q++) {
            mbar_arrive_expect_tx(full + grp * NG + q, (uint32_t)(kSpectrum * 16));
            bulk_copy_g2s(keys + ((size_t)grp * NG + q) * kSpectrum, key_ptr(i, q) - t, kSpectrum * 16, full + grp * NG + q);
        }
    };
    if (!KT && threadIdx.x < NO * NG) mbar_init(full + threadIdx.x, 1);
    if (!KT && threadIdx.x == 0) mbar_fence_init();
    if (warp == 0) tmem_alloc<kTmemCols>(&s_tmem_base);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const uint32_t tm = s_tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
    const uint32_t tm_own = tm + (uint32_t)((grp >> 1) * 32), tm_peer = tm + (uint32_t)(((grp >> 1) ^ 1) * 32);
    const uint32_t tm_key = tm + 64u + (uint32_t)((grp >> 1) * 96);
    if (!KT && outg && t == 0) issue_keys(0);

    double2 kq[8];   // KT = 1: the key spectrum in flight from L2
    auto key_load = [&](int i, int q) {
        const double2* src = key_ptr(i, q);
#pragma unroll
        for (int e = 0; e < 8; e++)   // volatile: issued HERE, a phase ahead of its use, not sunk to the use by the compiler
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(kq[e].x), "=d"(kq[e].y) : "l"(src + e * 64));
    };
    if (KT && outg) {
        key_load(0, 0);
        tmem_store_spectrum(tm_key, kq);
        key_load(0, 1);
    }

    lowlat_prologue(A, g, acc, bara);
    __syncthreads();

    const int c1 = grp / L, r1 = grp % L;   // phase-1 role
    const int32_t* p = acc + c1 * kN;
    long long pr[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ck = 0;   // PROBE: cycles per phase (rotate, forward, A, MAC, B, inverse, update, C)
    auto lap = [&](int k) { if (PROBE) { const long long n = clock64(); pr[k] += n - ck; ck = n; } };
#pragma unroll 1
    for (int i = 0; i < A.n_iter; i++) {   // bootstrap.jl:19-23; a zero rotation is executed (exact no-op)
        const int s = bara[i] & 2047;
        if (PROBE) ck = clock64();
        double2 a[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int j = t + 64 * m;
            const uint32_t tl = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;              // bootstrap.jl:21
            const uint32_t th = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
            a[m] = make_double2(digit_f64<BGBIT>(tl, r1), -digit_f64<BGBIT>(th, r1));                // tgsw.jl:104-116
        }
        lap(0);
        if (KT && outg) { tmem_store_spectrum(tm_key + 32, kq); key_load(i, 2); }   // q = 1 parked, q = 2 in flight behind the transform
        fft512_forward(a, w, X1, X2, t, bar_id);
        if (KT && outg) { tmem_store_spectrum(tm_key + 64, kq); key_load(i, 3); }   // q = 3 stays in registers
#pragma unroll
        for (int e = 0; e < 8; e++) X1[e * 64 + t] = a[e];   // X1 is free: every thread of the group passed the 2nd barrier
        tmem_store_spectrum(tm_own, a);
        tmem_wait_st();
        tmem_fence_before_sync();
        lap(1);
        __syncthreads();   // A: all spectra published, all reads of acc done
        tmem_fence_after_sync();
        lap(2);
        double2 o[8];
        if (outg) {
            int pr0[16], pr1[16], kr0[16], kr1[16];
            double2 P[8];
            tmem_ld_spectrum_raw(tm_peer, pr0, pr1);
            if (KT) tmem_ld_spectrum_raw(tm_key, kr0, kr1);
            tmem_ld_spectrum_finish(pr0, pr1, P);
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < NG; q++) {   // same summation order as K3: c outer, r inner
                double2 K[8];
                if (KT) {
                    if (q < 3) {
                        if (q == 0) {   // landed with the peer spectrum (one wait covers every outstanding load)
#pragma unroll
                            for (int x = 0; x < 16; x++) asm volatile("" : "+r"(kr0[x]), "+r"(kr1[x]));
#pragma unroll
                            for (int x = 0; x < 4; x++) {
                                K[x] = make_double2(__hiloint2double(kr0[4 * x + 1], kr0[4 * x]), __hiloint2double(kr0[4 * x + 3], kr0[4 * x + 2]));
                                K[4 + x] = make_double2(__hiloint2double(kr1[4 * x + 1], kr1[4 * x]), __hiloint2double(kr1[4 * x + 3], kr1[4 * x + 2]));
                            }
                        } else tmem_ld_spectrum_finish(kr0, kr1, K);
                        if (q < 2) tmem_ld_spectrum_raw(tm_key + (uint32_t)((q + 1) * 32), kr0, kr1);   // in flight behind this MAC
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; e++) K[e] = kq[e];
                    }
                } else {
                    mbar_wait(full + grp * NG + q, (uint32_t)i & 1u);
                    const double2* Ks = keys + ((size_t)grp * NG + q) * kSpectrum + t;
#pragma unroll
                    for (int e = 0; e < 8; e++) K[e] = Ks[e * 64];
                }
                if (q == grp) {
#pragma unroll
                    for (int e = 0; e < 8; e++) cmac(o[e], a[e], K[e]);                                  // tgsw.jl:128
                } else if (q == (grp ^ 2)) {
#pragma unroll
                    for (int e = 0; e < 8; e++) cmac(o[e], P[e], K[e]);
                } else {
                    const double2* F = xbuf + (size_t)q * (kSpectrum + kX2Elems) + t;
#pragma unroll
                    for (int e = 0; e < 8; e++) cmac(o[e], F[e * 64], K[e]);
                }
            }
        }
        tmem_fence_before_sync();
        lap(3);
        __syncthreads();   // B: spectra (shared memory and TMEM) and key slots consumed
        tmem_fence_after_sync();
        lap(4);
        if (outg) {
            const bool more = i + 1 < A.n_iter;
            if (!KT && t == 0 && more) issue_keys(i + 1);   // lands during the inverse + next forward transform
            if (KT && more) key_load(i + 1, 0);
            fft512_inverse(o, w, X1, X2, t, bar_id);
            if (KT && more) { tmem_store_spectrum(tm_key, kq); key_load(i + 1, 1); }
            lap(5);
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
        lap(6);
        __syncthreads();   // C: accumulator updated
        lap(7);
    }
    if (PROBE && A.probe && blockIdx.x == 0 && (threadIdx.x & 31) == 0)
        for (int k = 0; k < 8; k++) A.probe[warp * 8 + k] = (unsigned long long)pr[k];

    // tlwe_extract_sample (tlwe.jl:55-59): a = (p_0, -p_{N-1}, ..., -p_1), b = acc_b[0]
    int32_t* out = A.out + g * (kN + 1);
    for (int x = threadIdx.x; x < kN; x += blockDim.x) out[x] = x == 0 ? acc[0] : (int32_t)(0u - (uint32_t)acc[kN - x]);
    if (threadIdx.x == 0) out[kN] = acc[kN];
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<kTmemCols>(s_tmem_base);
}

}  // namespace tfhe_b200

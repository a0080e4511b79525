// blind_rotate_cluster.cuh — K3C: blind rotation of ONE gate by a CLUSTER of two CTAs (two SMs), two-piece transform
// (described for l = 2; l = 3 has three groups per CTA and six key spectra per output).
//
// K3L (blind_rotate_lowlat.cuh) runs the four forward and four inverse transforms of an iteration at once on one SM; both
// phases are bound by that SM's FP64 rate (clock64 probe).  Here CTA c of the pair owns accumulator
// component c for the whole blind rotation:
//
//   phase 1  group r of CTA c: rotate/subtract acc[c] (bootstrap.jl:21), digit r (tgsw.jl:104-116), forward transform;
//            the spectrum F_(c,r) goes to the CTA's own shared memory AND, with st.async (16 B per store, completing
//            bytes on an mbarrier of the peer), into a landing buffer of the other CTA         -> CTA barrier A
//   phase 2  group pc of CTA c owns output (c' = c, piece pc): first the two products with the CTA's own spectra
//            F_(c,0), F_(c,1) — they run while the peer's 16 KB are in flight (DSMEM moves ~11 B/clk each way,
//            tools/dsmem_test.cu) —, then, after the landing barrier, the two with the peer's        (tgsw.jl:128)
//                                                                                              -> CTA barrier B
//   phase 3  inverse transform, round (polynomials.jl:115-116), high piece << 16, integer atomic add into acc[c]
//                                                                                              -> CTA barrier C
//
// Key: the four spectra BK[i][r][cq][c'=c][pc] of an output are copied (TMA bulk copies, one mbarrier per group) into one of
// TWO sets of slots, the set of iteration i+1 right after barrier A of iteration i: a 64 KB fill shares the shared-memory
// pipe with whatever runs beside it and was measured to lengthen a transform phase by ~700 cycles, while in the window
// after barrier A the warps sit behind their remote stores anyway (clock64 probe: 1.276 -> 1.111 ms).  Reading the key
// straight into registers instead (coalesced 16-byte loads, a phase ahead) was slower (1.38 ms).
//
// acc[c] never leaves its CTA; the only traffic between the SMs is the spectra, and the only cluster-wide
// synchronisation is their landing barrier (two landing buffers, used alternately: a CTA can only send the spectra of
// iteration i+2 after it has consumed the peer's of iteration i+1, which the peer sent after it had finished reading
// iteration i).  The sum over q is taken own-spectra-first, i.e. in the order (2,3,0,1) in CTA 1: this kernel is only
// used with the two-piece transform, where every partial sum is an exact integer below 2^53 by the bound of DESIGN.md §4,
// so the rounded result does not depend on the order and the output is bit-identical to K3 / K3L (tests).
#pragma once
#include "blind_rotate_lowlat.cuh"

namespace tfhe_b200 {

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_map(uint32_t smem_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes into the peer's shared memory; the peer's mbarrier counts them
__device__ __forceinline__ void st_async16(uint32_t remote_addr, double2 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                 ::"r"(remote_addr), "l"(__double_as_longlong(v.x)), "l"(__double_as_longlong(v.y)), "r"(remote_bar) : "memory");
}

// l = 2: all four key spectra of an output fit twice (two sets of slots, filled alternately); l = 3 (128-bit set): six
// spectra per output, one set (96 KB) beside 48 KB of landing buffers, refilled after barrier B
template <int L> __host__ __device__ constexpr int br_cluster_key_sets() { return L == 2 ? 2 : 1; }
// l = 3: a seventh warp that only issues the key copies (1.83 -> 1.70 ms); l = 2: the first warp of each output group issues
// them itself (with the extra warp: 1.074 instead of 1.062 ms — it slows the transforms of its sub-partition)
template <int L> __host__ __device__ constexpr int br_cluster_threads() { return 64 * L + (L == 3 ? 32 : 0); }
template <int L> __host__ __device__ constexpr size_t br_cluster_smem_bytes(int n_pad) {
    return (size_t)br_cluster_key_sets<L>() * 2 * (2 * L) * kSpectrum * 16   // key slots [set][output group][q]
           + (size_t)2 * L * kSpectrum * 16           // landing buffers [2][r]
           + 128                                      // mbarriers: sets x 2 key + 2 landing
           + (size_t)L * (kSpectrum + kX2Elems) * 16  // X1, X2 per group
           + 2 * kN * 4 + (size_t)n_pad * 4;          // accumulator (both components initialised, one maintained), mask
}

// L groups of 64 threads per CTA: group r transforms digit r; groups 0 and 1 also own the outputs (c, low piece) and
// (c, high piece).
template <int L, int BGBIT, int PROBE = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(br_cluster_threads<L>(), 1) blind_rotate_cluster_kernel(BlindRotateArgs A) {
    constexpr int NP = 2, NQ = 2 * L, KSETS = br_cluster_key_sets<L>();
    constexpr int NT = 64 * L, NTP = br_cluster_threads<L>();   // compute threads; + the warp that only issues the key copies (l = 3)
    constexpr bool KW = NTP > NT;
    // CTA barriers as named barriers: the key warp joins only the one after which it issues (A with two sets of slots, B with one)
    auto bar_a = [&]() { if (KSETS == 2) asm volatile("bar.sync 8, %0;" ::"n"(NTP) : "memory"); else asm volatile("bar.sync 8, %0;" ::"n"(NT) : "memory"); };
    auto bar_b = [&]() { if (KSETS == 1) asm volatile("bar.sync 9, %0;" ::"n"(NTP) : "memory"); else asm volatile("bar.sync 9, %0;" ::"n"(NT) : "memory"); };
    auto bar_c = [&]() { asm volatile("bar.sync 10, %0;" ::"n"(NT) : "memory"); };
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* keys = reinterpret_cast<double2*>(smem_raw);                         // [set][output group][q][512]
    double2* land = keys + (size_t)KSETS * 2 * NQ * kSpectrum;                    // [buf][r][512]
    uint64_t* kbar = reinterpret_cast<uint64_t*>(land + (size_t)2 * L * kSpectrum);   // [set][output group]
    uint64_t* lbar = kbar + 2 * KSETS;                                            // [buf][r]: one per landing spectrum
    double2* xbuf = reinterpret_cast<double2*>(reinterpret_cast<unsigned char*>(kbar) + 128);
    int32_t* acc = reinterpret_cast<int32_t*>(xbuf + (size_t)L * (kSpectrum + kX2Elems));
    int32_t* bara = acc + 2 * kN;

    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int bar_id = grp + 1;
    const int c = (int)cluster_ctarank();      // accumulator component of this CTA = output component c'
    const int peer = c ^ 1;
    const bool outg = grp < NP;                // this group also owns output (c, piece grp)
    double2* X1 = xbuf + (size_t)grp * (kSpectrum + kX2Elems);
    double2* X2 = X1 + kSpectrum;
    const unsigned long long g = blockIdx.x >> 1;
    // 128 threads leave room for the full twiddle set; 192 threads (168 registers) keep the compact one
    typename std::conditional<L == 2, TwiddlesFull, Twiddles>::type w; w.load(A.E, t);

    // output group pc reads, for q = (cq, r), the spectrum BK[i][r][cq][c][pc].  One barrier per (slot set, group) counts all
    // copies of an iteration; one lane per copy issues them in the same instructions (a warp that issues a bulk copy loses
    // ~350-450 cycles whatever the number of copies).
    auto issue_keys = [&](int i) {   // called by one whole warp: the key warp (all outputs) or the first warp of an output group (its own)
        const int lane = threadIdx.x & 31, kg = KW ? lane / NQ : grp, q = lane % NQ;
        const int ks = KSETS == 2 ? (i & 1) : 0;
        const bool mine = lane < (KW ? 2 * NQ : NQ);
        if (mine && q == 0) mbar_arrive_expect_tx(kbar + ks * 2 + kg, (uint32_t)(NQ * kSpectrum * 16));
        __syncwarp();
        if (mine) {
            const int cq = q / L, r = q % L;
            const double2* src = A.bk_fft + ((((size_t)i * L + r) * 2 + cq) * 2 * NP + (size_t)c * NP + kg) * kSpectrum;
            bulk_copy_g2s(keys + ((size_t)(ks * 2 + kg) * NQ + q) * kSpectrum, src, kSpectrum * 16, kbar + ks * 2 + kg);
        }
    };
    const bool issuer = KW ? threadIdx.x >= NT : (outg && t < 32);
    if (threadIdx.x < 2 * KSETS + 2 * L) mbar_init(kbar + threadIdx.x, 1);   // key + landing barriers, contiguous
    if (threadIdx.x == 0) mbar_fence_init();
    __syncthreads();
    if (issuer) issue_keys(0);
    lowlat_prologue(A, g, acc, bara);
    cluster_sync_all();   // both CTAs' barriers are initialised before the first remote store can arrive (also a CTA barrier)

    if (KW && threadIdx.x >= NT) {   // the key warp
        for (int i = 0; i < A.n_iter; i++) {
            if (KSETS == 2) bar_a(); else bar_b();
            if (i + 1 < A.n_iter) issue_keys(i + 1);
        }
        cluster_sync_all();
        return;
    }
    const uint32_t r_land = cluster_map(smem_u32(land), (uint32_t)peer), r_lbar = cluster_map(smem_u32(lbar), (uint32_t)peer);
    int32_t* p = acc + c * kN;
    // PROBE: cycles per phase (rotate, forward, send + publish, barrier A, own products, landing wait, peer products,
    // barrier B, inverse, update + barrier C)
    long long pr[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, ck = 0;
    auto lap = [&](int k) { if (PROBE) { const long long n = clock64(); pr[k] += n - ck; ck = n; } };
#pragma unroll 1
    for (int i = 0; i < A.n_iter; i++) {   // bootstrap.jl:19-23; a zero rotation is executed (exact no-op)
        const int s = bara[i] & 2047;
        const int buf = i & 1;
        const int ks = KSETS == 2 ? buf : 0;
        const uint32_t kpar = (uint32_t)(KSETS == 2 ? (i >> 1) : i) & 1u;
        if (PROBE) ck = clock64();
        if (threadIdx.x < L) mbar_arrive_expect_tx(lbar + buf * L + threadIdx.x, (uint32_t)(kSpectrum * 16));   // the peer's spectra, one barrier each
        double2 a[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int j = t + 64 * m;
            const uint32_t tl = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;              // bootstrap.jl:21
            const uint32_t th = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
            a[m] = make_double2(digit_f64<BGBIT>(tl, grp), -digit_f64<BGBIT>(th, grp));              // tgsw.jl:104-116
        }
        lap(0);
        fft512_forward(a, w, X1, X2, t, bar_id);
        lap(1);
        // Group 1 sends AFTER barrier A and its first product: the fabric moves ~12 B/clk whatever the number of senders, so
        // staggering the sends lets the peer's first spectrum land ~half a transfer earlier, and its product (one landing
        // barrier per spectrum) runs while the second one is still in flight (1.09 -> 1.05 ms)
        const bool late_send = L == 2 && grp == 1;
        if (!late_send) {
            const uint32_t dst = r_land + (uint32_t)(((buf * L + grp) * kSpectrum + t) * 16);
#pragma unroll
            for (int e = 0; e < 8; e++) st_async16(dst + (uint32_t)(e * 64 * 16), a[e], r_lbar + (uint32_t)((buf * L + grp) * 8));
        }
#pragma unroll
        for (int e = 0; e < 8; e++) X1[e * 64 + t] = a[e];   // X1 is free: every thread of the group passed the 2nd barrier
        lap(2);
        // A: the sibling groups' spectra are published, all reads of acc done.  With two sets of slots the key warp now
        // issues the next iteration's key into the other set — the fill (64 KB through the shared-memory pipe) then
        // coincides with the window in which the warps sit behind their remote stores anyway, not with a transform
        // (issued after barrier B it made the inverse transform 700 cycles longer)
        bar_a();
        lap(3);
        if (!KW && KSETS == 2 && issuer && i + 1 < A.n_iter) issue_keys(i + 1);
        double2 o[8];
        if (outg) {
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = make_double2(0.0, 0.0);
            mbar_wait(kbar + ks * 2 + grp, kpar);   // the key spectra of this iteration (issued an iteration ago)
            {   // this group's own spectrum straight from registers ...
                const double2* K = keys + ((size_t)(ks * 2 + grp) * NQ + c * L + grp) * kSpectrum + t;
#pragma unroll
                for (int e = 0; e < 8; e++) cmac(o[e], a[e], K[e * 64]);                              // tgsw.jl:128
            }
            if (late_send) {
                const uint32_t dst = r_land + (uint32_t)(((buf * L + grp) * kSpectrum + t) * 16);
#pragma unroll
                for (int e = 0; e < 8; e++) st_async16(dst + (uint32_t)(e * 64 * 16), a[e], r_lbar + (uint32_t)((buf * L + grp) * 8));
            }
#pragma unroll
            for (int d = 1; d < L; d++) {   // ... then the sibling groups', while the peer's are in flight
                const int r = grp + d < L ? grp + d : grp + d - L;
                const double2* F = xbuf + (size_t)r * (kSpectrum + kX2Elems) + t;
                const double2* K = keys + ((size_t)(ks * 2 + grp) * NQ + c * L + r) * kSpectrum + t;
#pragma unroll
                for (int e = 0; e < 8; e++) cmac(o[e], F[e * 64], K[e * 64]);
            }
        }
        lap(4);
        lap(5);
        if (outg) {
#pragma unroll
            for (int r = 0; r < L; r++) {
                mbar_wait(lbar + buf * L + r, (uint32_t)(i >> 1) & 1u);   // the peer's spectrum r has landed
                const double2* F = land + (size_t)(buf * L + r) * kSpectrum + t;
                const double2* K = keys + ((size_t)(ks * 2 + grp) * NQ + peer * L + r) * kSpectrum + t;
#pragma unroll
                for (int e = 0; e < 8; e++) cmac(o[e], F[e * 64], K[e * 64]);
            }
        }
        lap(6);
        bar_b();   // B: published spectra and key slots consumed (one set of slots: refilled now)
        lap(7);
        if (!KW && KSETS == 1 && issuer && i + 1 < A.n_iter) issue_keys(i + 1);
        if (outg) {
            fft512_inverse(o, w, X1, X2, t, bar_id);
            lap(8);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                uint32_t vl = round_to_u32_fast<true>(o[m].x), vh = round_to_u32_fast<true>(-o[m].y);   // polynomials.jl:115-116
                if (grp == 1) { vl <<= 16; vh <<= 16; }
                const int j = t + 64 * m;
                atomicAdd(reinterpret_cast<unsigned int*>(p + j), vl);                                   // bootstrap.jl:22
                atomicAdd(reinterpret_cast<unsigned int*>(p + j + 512), vh);
            }
        }
        bar_c();   // C: accumulator updated
        lap(9);
    }
    if (PROBE && A.probe && blockIdx.x < 2 && (threadIdx.x & 31) == 0 && grp < 2)
        for (int k = 0; k < 10; k++) A.probe[(blockIdx.x * 4 + (threadIdx.x >> 5)) * 10 + k] = (unsigned long long)pr[k];

    // tlwe_extract_sample (tlwe.jl:55-59): a = (p_0, -p_{N-1}, ..., -p_1) from component 0, b = acc_b[0] from component 1
    int32_t* out = A.out + g * (kN + 1);
    if (c == 0) {
        for (int x = threadIdx.x; x < kN; x += NT) out[x] = x == 0 ? p[0] : (int32_t)(0u - (uint32_t)p[kN - x]);
    } else if (threadIdx.x == 0) out[kN] = p[0];
    cluster_sync_all();   // neither CTA leaves while the other could still address its shared memory
}

}  // namespace tfhe_b200

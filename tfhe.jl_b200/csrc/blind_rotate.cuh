// blind_rotate.cuh — K3: the persistent blind-rotation kernel (bootstrap.jl:19-82) for sm_100a.
//
// One CTA keeps G gates resident for all n iterations: one 64-thread group per gate, the TLWE accumulator
// of every gate in shared memory.  The G groups walk the bootstrapping key in lockstep, so each 16 KB key
// chunk is fetched from L2 ONCE per CTA by a TMA bulk copy (cp.async.bulk, SASS UBLKCP) into a ring of
// shared-memory stages and then read by all G gates with conflict-free LDS.128:
//
//   producer (thread 0 of the CTA):  wait empty[s] -> arrive.expect_tx(full[s], 16 KB) -> cp.async.bulk,
//                                    STAGES-1 chunks ahead of its own consumption
//   consumer groups (2 warps each):  forward FFT of a digit polynomial (no key needed) ->
//                                    wait full[s] -> multiply-accumulate against the chunk -> arrive empty[s]
//
// The copy of chunk k+1.. overlaps the transform of chunk k, so L2 latency never reaches the FP64 pipe
// (v1 read the key with ld.global.nc inside the MAC: ncu showed 44 % of all stall samples on those DFMAs
// waiting on the long scoreboard, profiles/r1/).  L2->SM key traffic drops by G x.
//
// bootstrap.jl:34 skips iterations whose rotation is 0.  Lockstep groups cannot skip independently, so a
// zero rotation is simply executed: temp = X^0*acc - acc = 0, all digits are 0 (tgsw.jl:99-117 maps 0 to 0),
// the products are exactly 0 and the accumulator is unchanged — bit-identical to skipping.
#pragma once
#include "kernels.cuh"
#include "tmem.cuh"
#include <type_traits>

namespace tfhe_b200 {

// ---- mbarrier / bulk-copy primitives (PTX) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// The same wait without the hardware suspend: mbarrier.try_wait parks the warp until the phase completes OR a
// system-dependent time limit expires; a warp that starts waiting shortly before a bulk copy lands was measured to
// lose several hundred cycles that way (clock64 probe).  test_wait only polls.
__device__ __forceinline__ void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (TMA, no tensor map needed for a contiguous 1-D block), completion on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction policy for the lines it touches
__device__ __forceinline__ void bulk_copy_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}

constexpr int kChunkElems = 2 * kSpectrum;        // double2 per key chunk
constexpr int kChunkBytes = kChunkElems * 16;     // 16 KB

// View of the key ring.  Every consumer thread tracks (stage, phase); thread 0 of the CTA is also the
// producer: just before it waits for chunk k it refills the stage that chunk k-1 occupied with chunk
// k+STAGES-1 (all groups released it an FFT ago, so the wait on `empty` practically never blocks), which
// keeps STAGES-1 chunks in flight without spending a warp — a 9th warp would put 3 warps on one SM
// sub-partition and cap every thread at 168 registers (16K registers per sub-partition).
// ORDER == 0: chunks are consumed in the order (c, r, half); ORDER == 1 (two-piece transforms, output-stationary
// step): (c', c, r) — all digit polynomials against the two pieces of output component c' = 0, then c' = 1.
//
// PW == 0: thread 0 of the CTA is also the producer: just before it waits for chunk k it refills the stage that chunk
//          k-1 occupied with chunk k+STAGES-1.
// PW == 1: a dedicated warp walks the ring (produce_all) and the compute warps only consume; every consumer also tests
//          the NEXT chunk's barrier one chunk early.  Measured with the clock64 probe: an mbarrier test
//          costs 100-150 cycles even when the phase is long complete; with the in-line producer that latency (twice per
//          chunk: empty + full) plus the issue code made warp 0 the straggler of the CTA (8.2 k cycles in the
//          multiply-accumulate phase against 5.9 k for its sibling warp) and every other group waited for it.
template <int L, int NP, int STAGES, int ORDER = 0, int PROBE = 0, int PW = 0> struct BkFromRing {
    static constexpr int kChunksPerIter = 2 * L * NP;
    const double2* ring; uint64_t* full; uint64_t* empty;
    const double2* bk;        // start of the key, [n_iter][kChunksPerIter] chunks
    int total;                // chunks in the whole walk
    int stage; uint32_t phase;
    bool producer;            // PW == 0: thread 0 of the CTA
    long long w_full = 0, w_empty = 0, w_first = 0;   // PROBE: cycles spent waiting (all chunks / producer / first chunk of a pass)
    // producer cursor: the next chunk to issue (sequence number, its stage, how often that stage has been used, its
    // position (iteration, index in the iteration)) — advanced incrementally, no division in the loop
    int iss = 0, iss_stage = 0, iss_round = 0, iss_i = 0, iss_q = 0;
    uint32_t ready_next = 0;  // PW: result of the early test of the next chunk's full barrier
    uint64_t policy = 0;      // != 0: L2 eviction policy of the key lines (the key outlives the ciphertext stream in L2)

    // index in the iteration (consumption order) -> chunk inside the stored row: storage order is (r, c, half)
    static __device__ __forceinline__ int stored_index(int q) {
        int h, cr;
        if (ORDER == 0) { h = q % NP; cr = q / NP; } else { h = q / (2 * L); cr = q % (2 * L); }
        const int r = cr % L, c = cr / L;
        return (r * 2 + c) * NP + h;
    }
    __device__ __forceinline__ void issue_next() {
        const size_t off = ((size_t)iss_i * kChunksPerIter + (size_t)stored_index(iss_q)) * kChunkElems;
        mbar_arrive_expect_tx(full + iss_stage, kChunkBytes);
        if (policy) bulk_copy_g2s_hint(const_cast<double2*>(ring) + (size_t)iss_stage * kChunkElems, bk + off, kChunkBytes, full + iss_stage, policy);
        else bulk_copy_g2s(const_cast<double2*>(ring) + (size_t)iss_stage * kChunkElems, bk + off, kChunkBytes, full + iss_stage);
        iss++;
        if (++iss_stage == STAGES) { iss_stage = 0; iss_round++; }
        if (++iss_q == kChunksPerIter) { iss_q = 0; iss_i++; }
    }
    __device__ __forceinline__ void prologue() {
        if (!PW && producer)
            while (iss < STAGES - 1 && iss < total) issue_next();
    }
    // PW: the whole walk, by one thread of a warp that does nothing else — it runs as far ahead of the consumers as the
    // ring is deep and only ever waits for a stage to drain, never for data it needs itself
    __device__ __forceinline__ void produce_all() {
        while (iss < total) {
            if (iss_round >= 1) mbar_wait(empty + iss_stage, (uint32_t)(iss_round - 1) & 1u);
            issue_next();
        }
    }
    static __device__ __forceinline__ uint32_t test(uint64_t* bar, uint32_t parity) {
        uint32_t done;
        asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        return done;
    }
    __device__ __forceinline__ const double2* acquire(int first_of_pass /*consumption order is fixed*/) {
        if (!PW && producer && iss < total) {
            long long c0 = 0;
            if (PROBE) c0 = clock64();
            // the stage's previous occupant (chunk iss - STAGES) must have been released by every warp
            if (iss_round >= 1) mbar_wait(empty + iss_stage, (uint32_t)(iss_round - 1) & 1u);
            if (PROBE) w_empty += clock64() - c0;
            issue_next();
        }
        long long c1 = 0;
        if (PROBE) c1 = clock64();
        if (PW) {
            if (!ready_next) mbar_wait_poll(full + stage, phase);
            // test the following chunk now; the answer is back long before the next acquire needs it
            const int ns = stage + 1 == STAGES ? 0 : stage + 1;
            ready_next = test(full + ns, ns == 0 ? phase ^ 1u : phase);
        } else {
            mbar_wait(full + stage, phase);
        }
        if (PROBE) { const long long d = clock64() - c1; w_full += d; if (first_of_pass) w_first += d; }
        return ring + (size_t)stage * kChunkElems;
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    static __device__ __forceinline__ double2 load(const double2* p) { return *p; }
};

// K2 with the (k+1)*NP output spectra held in TENSOR MEMORY instead of registers (tmem.cuh).  The register
// version (kernels.cuh) needs 64 / 128 registers of accumulators per thread, which limits K3 to 8 warps per
// SM; with the accumulators in TMEM a thread needs < 168 registers, 12 warps (6 gates) fit, and the
// read-modify-write of the accumulators runs on the TMEM datapath (LDTM/STTM), off the shared-memory pipe.
// `tm` is the TMEM address of this thread's row (lane and first column already applied); the accumulator of
// output component c2, piece pc occupies columns [(c2*NP + pc)*32, +32).
// REGH = 1 (NP == 2 only): the two pieces of output component 0 stay in registers (64 registers) and only
// component 1 goes through TMEM, which halves the TMEM round trips.
template <int L, int BGBIT, int NP, int REGH, class BK>
__device__ __forceinline__ void extern_product_step_tmem(int32_t* acc, int abar, BK& bk, const Twiddles& w, double2* X1,
                                                         double2* X2, uint32_t tm, int t, int bar_id) {
    static_assert(REGH == 0 || NP == 2, "REGH needs the two-piece transform");
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
    double2 oreg[REGH ? 2 : 1][8];
#pragma unroll
    for (int sp = 0; sp < (REGH ? 2 : 1); sp++)
#pragma unroll
        for (int q = 0; q < 8; q++) oreg[sp][q] = make_double2(0.0, 0.0);
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        const int32_t* p = acc + c * kN;
        uint32_t tl[8], th[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            tl[m] = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;                 // bootstrap.jl:21
            th[m] = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
        }
#pragma unroll 1
        for (int r = 0; r < L; r++) {
            double2 a[8];
#pragma unroll
            for (int m = 0; m < 8; m++)
                a[m] = make_double2(digit_f64<BGBIT>(tl[m], r), -digit_f64<BGBIT>(th[m], r));   // tgsw.jl:104-116
            double2 kv[REGH ? 16 : 1];
            fft512_forward(a, w, X1, X2, t, bar_id, [&]() {
                if (REGH) {   // first key chunk of (r, c) loaded behind the transform's last exchange
                    const double2* b = bk.acquire((r * 2 + c) * NP) + t;
#pragma unroll
                    for (int e = 0; e < 16; e++) kv[e % (REGH ? 16 : 1)] = BK::load(b + e * 64);
                }
            });
            const bool first = (c == 0 && r == 0);
            if (!first) tmem_wait_st();   // this thread's previous accumulator stores have landed
            if (REGH) {
#pragma unroll
                for (int sp = 0; sp < 2; sp++)
#pragma unroll
                    for (int q = 0; q < 8; q++) cmac(oreg[sp][q], a[q], kv[(sp * 8 + q) % (REGH ? 16 : 1)]);
                bk.release();
            }
#pragma unroll
            for (int half = REGH ? 1 : 0; half < NP; half++) {
                const double2* b = bk.acquire((r * 2 + c) * NP + half) + t;
#pragma unroll
                for (int sp = 0; sp < 2; sp++) {
                    const uint32_t col = tm + (uint32_t)((NP == 1 ? sp : half * 2 + sp) * 32);
                    double2 o[8];
                    if (first) {
#pragma unroll
                        for (int q = 0; q < 8; q++) o[q] = make_double2(0.0, 0.0);
                    } else {
                        tmem_load_spectrum(col, o);
                    }
#pragma unroll
                    for (int q = 0; q < 8; q++) cmac(o[q], a[q], BK::load(b + (sp * 8 + q) * 64));      // tgsw.jl:128
                    tmem_store_spectrum(col, o);
                }
                bk.release();
            }
        }
    }
    tmem_wait_st();
    group_sync(bar_id);   // all reads of acc and of X2 (last forward) are done
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++) {
        uint32_t rl[8], rh[8];
#pragma unroll
        for (int pc = 0; pc < NP; pc++) {
            double2 o[8];
            if (REGH && c2 == 0) {
#pragma unroll
                for (int q = 0; q < 8; q++) o[q] = oreg[pc % 2][q];
            } else {
                tmem_load_spectrum(tm + (uint32_t)((c2 * NP + pc) * 32), o);
            }
            fft512_inverse(o, w, X1, X2, t, bar_id);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                uint32_t vl = round_to_u32_fast<NP == 2>(o[m].x), vh = round_to_u32_fast<NP == 2>(-o[m].y);                      // polynomials.jl:115-116
                if (pc == 0) { rl[m] = vl; rh[m] = vh; }
                else { rl[m] += vl << 16; rh[m] += vh << 16; }
            }
        }
        int32_t* p = acc + c2 * kN;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            p[j] = (int32_t)((uint32_t)p[j] + rl[m]);                                               // bootstrap.jl:22
            p[j + 512] = (int32_t)((uint32_t)p[j + 512] + rh[m]);
        }
    }
    group_sync(bar_id);
}

// K2, output-stationary form for the two-piece (proven-exact) transform.
//
// extern_product_step_tmem keeps 2*NP = 4 output spectra alive while the 2L digit polynomials stream through
// (64 accumulator registers + 64 TMEM columns that are read, updated and written back for every digit polynomial).
// Here the loop nest is turned inside out:
//
//   phase 1   for q = (c, r):  rotate/subtract, digit r, forward transform  ->  F_q parked in TENSOR MEMORY
//             (2L x 32 columns of this thread's TMEM row, written once)
//   phase 2   for c' = 0, 1:   (lo, hi) = sum_q F_q * (BK[r][c][c'][piece 0], [piece 1])   one 16 KB ring chunk per q
//                              inverse transform lo, hi; round; acc[c'] += lo + (hi << 16)
//
// so an output spectrum lives in registers from its first multiply-accumulate to its inverse transform, TMEM is
// written once and read twice per F_q with no read-modify-write dependency (the load of F_{q+1} is in flight
// behind the MAC of F_q), and the ~100 registers this frees hold all 15 twiddles of the thread (TwiddlesFull:
// no twiddle is re-derived, 44 FP64 instructions less per transform).  With SYNC == 1 transforms the group meets at
// one barrier per transform + one per iteration (9 instead of 19).
//
// X1 points at TWO consecutive 512-element buffers when SYNC == 1 (see fft512_forward_t).
// The sums over q are taken in the same order as in every other kernel (c outer, r inner); in this mode all
// results are exact integers anyway (exactness).
__device__ __forceinline__ void tmem_ld_spectrum_raw(uint32_t taddr, int (&r0)[16], int (&r1)[16]) {
    tmem_ld4_raw(taddr, r0);
    tmem_ld4_raw(taddr + 16, r1);
}
// wait for the loads and pin the register reads behind the wait (an empty volatile asm that "modifies" them)
__device__ __forceinline__ void tmem_ld_spectrum_finish(int (&r0)[16], int (&r1)[16], double2 (&v)[8]) {
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; i++) asm volatile("" : "+r"(r0[i]), "+r"(r1[i]));
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[i] = make_double2(__hiloint2double(r0[4 * i + 1], r0[4 * i]), __hiloint2double(r0[4 * i + 3], r0[4 * i + 2]));
        v[4 + i] = make_double2(__hiloint2double(r1[4 * i + 1], r1[4 * i]), __hiloint2double(r1[4 * i + 3], r1[4 * i + 2]));
    }
}

// The same iteration with the transforms taken in pairs (fft512_forward_dual / _inverse_dual): the two digit polynomials
// of an accumulator component (l = 2), and the two pieces of an output component.
template <int BGBIT, class BK, class W>
__device__ __forceinline__ void extern_product_step_os_dual(int32_t* acc, int abar, BK& bk, const W& w, double2* X1, double2* X2,
                                                            uint32_t tm, int t, int bar_id) {
    constexpr int L = 2, Q = 4;
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        const int32_t* p = acc + c * kN;
        double2 a[8], b[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int j = t + 64 * m;
            const uint32_t tl = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;                 // bootstrap.jl:21
            const uint32_t th = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
            a[m] = make_double2(digit_f64<BGBIT>(tl, 0), -digit_f64<BGBIT>(th, 0));                     // tgsw.jl:104-116
            b[m] = make_double2(digit_f64<BGBIT>(tl, 1), -digit_f64<BGBIT>(th, 1));
        }
        fft512_forward_dual(a, b, w, X1, X1 + kSpectrum, X2, t, bar_id);
        tmem_store_spectrum(tm + (uint32_t)((c * L) * 32), a);
        tmem_store_spectrum(tm + (uint32_t)((c * L + 1) * 32), b);
    }
    tmem_wait_st();
#pragma unroll 1
    for (int c2 = 0; c2 < 2; c2++) {
        double2 lo[8], hi[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { lo[e] = make_double2(0.0, 0.0); hi[e] = make_double2(0.0, 0.0); }
        int r0[16], r1[16];
        tmem_ld_spectrum_raw(tm, r0, r1);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            double2 F[8];
            tmem_ld_spectrum_finish(r0, r1, F);
            if (q + 1 < Q) tmem_ld_spectrum_raw(tm + (uint32_t)((q + 1) * 32), r0, r1);
            const double2* kb = bk.acquire(q == 0) + t;
#pragma unroll
            for (int e = 0; e < 8; e++) cmac(lo[e], F[e], BK::load(kb + e * 64));             // tgsw.jl:128
#pragma unroll
            for (int e = 0; e < 8; e++) cmac(hi[e], F[e], BK::load(kb + (8 + e) * 64));
            bk.release();
        }
        fft512_inverse_dual(lo, hi, w, X1, X1 + kSpectrum, X2, t, bar_id);
        int32_t* p = acc + c2 * kN;
#pragma unroll
        for (int m = 0; m < 8; m++) {                                                          // polynomials.jl:115-116, bootstrap.jl:22
            const int j = t + 64 * m;
            p[j] = (int32_t)((uint32_t)p[j] + round_to_u32_fast<true>(lo[m].x) + (round_to_u32_fast<true>(hi[m].x) << 16));
            p[j + 512] = (int32_t)((uint32_t)p[j + 512] + round_to_u32_fast<true>(-lo[m].y) + (round_to_u32_fast<true>(-hi[m].y) << 16));
        }
    }
    group_sync(bar_id);   // the updated accumulator is visible to the whole group before the next rotation reads it
}

template <int L, int BGBIT, int SYNC, int PROBE, class BK, class W>
__device__ __forceinline__ void extern_product_step_os(int32_t* acc, int abar, BK& bk, const W& w, double2* X1, double2* X2,
                                                       uint32_t tm, int t, int bar_id, long long (&pr)[4]) {
    constexpr int Q = 2 * L;
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
    long long c0 = 0, c1 = 0;
    if (PROBE) c0 = clock64();
    // ---- phase 1: the Q forward transforms ----
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
        const int32_t* p = acc + c * kN;
        uint32_t tl[8], th[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            tl[m] = (uint32_t)rot_coeff(p, j, s) - (uint32_t)p[j] + offset;                 // bootstrap.jl:21
            th[m] = (uint32_t)rot_coeff(p, j + 512, s) - (uint32_t)p[j + 512] + offset;
        }
#pragma unroll 1
        for (int r = 0; r < L; r++) {
            double2 a[8];
#pragma unroll
            for (int m = 0; m < 8; m++)
                a[m] = make_double2(digit_f64<BGBIT>(tl[m], r), -digit_f64<BGBIT>(th[m], r));   // tgsw.jl:104-116
            const int q = c * L + r;
            fft512_forward_t<SYNC>(a, w, X1 + (SYNC ? (q & 1) * kSpectrum : 0), X2, t, bar_id, NoPrefetch());
            tmem_store_spectrum(tm + (uint32_t)(q * 32), a);
        }
    }
    tmem_wait_st();   // this thread's F_q have landed; every acc read of the group precedes the last transform's barrier
    if (PROBE) { c1 = clock64(); pr[0] += c1 - c0; }
    // ---- phase 2: one output component at a time ----
#pragma unroll 1
    for (int c2 = 0; c2 < 2; c2++) {
        if (PROBE) c0 = clock64();
        double2 lo[8], hi[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { lo[e] = make_double2(0.0, 0.0); hi[e] = make_double2(0.0, 0.0); }
        int r0[16], r1[16];
        tmem_ld_spectrum_raw(tm, r0, r1);
#pragma unroll
        for (int q = 0; q < Q; q++) {
            double2 F[8];
            tmem_ld_spectrum_finish(r0, r1, F);
            if (q + 1 < Q) tmem_ld_spectrum_raw(tm + (uint32_t)((q + 1) * 32), r0, r1);   // in flight behind the MAC below
            const double2* b = bk.acquire(q == 0) + t;
#pragma unroll
            for (int e = 0; e < 8; e++) cmac(lo[e], F[e], BK::load(b + e * 64));             // tgsw.jl:128
#pragma unroll
            for (int e = 0; e < 8; e++) cmac(hi[e], F[e], BK::load(b + (8 + e) * 64));
            bk.release();
        }
        if (PROBE) { c1 = clock64(); pr[1] += c1 - c0; }
        uint32_t rl[8], rh[8];
        fft512_inverse_t<SYNC>(lo, w, X1, X2, t, bar_id);
#pragma unroll
        for (int m = 0; m < 8; m++) { rl[m] = round_to_u32_fast<true>(lo[m].x); rh[m] = round_to_u32_fast<true>(-lo[m].y); }   // polynomials.jl:115-116
        fft512_inverse_t<SYNC>(hi, w, X1 + (SYNC ? kSpectrum : 0), X2, t, bar_id);
        int32_t* p = acc + c2 * kN;
#pragma unroll
        for (int m = 0; m < 8; m++) {
            const int j = t + 64 * m;
            p[j] = (int32_t)((uint32_t)p[j] + rl[m] + (round_to_u32_fast<true>(hi[m].x) << 16));                           // bootstrap.jl:22
            p[j + 512] = (int32_t)((uint32_t)p[j + 512] + rh[m] + (round_to_u32_fast<true>(-hi[m].y) << 16));
        }
        if (PROBE) pr[2] += clock64() - c1;
    }
    if (PROBE) c0 = clock64();
    group_sync(bar_id);   // the updated accumulator is visible to the whole group before the next rotation reads it
    if (PROBE) pr[3] += clock64() - c0;
}

// ---- K4T: key switch of a TILE of 64 ciphertexts per CTA (keyswitch.jl:45-80) ---------------------------------
// keyswitch_kernel (one CTA per ciphertext) gathers 12.3 MB of table rows per ciphertext from L2 and runs at the
// L2 bandwidth limit (13.9 TB/s).  Here a CTA owns 64 ciphertexts and streams the WHOLE table once through shared
// memory — the 24 rows of one mask position i (t = 8 digits x 3 non-zero values), double buffered by TMA bulk
// copies — so L2 traffic drops from 12.3 MB to 0.8 MB per ciphertext and the gather becomes LDS.128 reads of
// rows that every lane of a warp shares: a warp covers 128 output columns (lane = column quad) of 32 ciphertexts
// whose 32 x 4 partial sums live in registers.  Branch-free: a stage holds 4 slots per digit position, slot 0 a
// row of zeros that is never overwritten, so digit value d simply selects slot d.
// Integer subtraction mod 2^32 commutes: the result is the same bits as keyswitch_kernel.
constexpr int kKsTile = 64;      // ciphertexts per CTA
constexpr int kKsT = 8, kKsBasebit = 2;
__host__ __device__ inline size_t ks_tile_smem_bytes(int stride) { return 2 * (size_t)kKsT * 4 * stride * 4 + 64; }

// HALVES = 2: 64 ciphertexts per CTA.  HALVES = 1: 32 per CTA, half the warps — for batches whose 64-ciphertext tiles
// would not fill the SMs: a tile of 32 is done in about half the time, so the fixed cost of a launch drops with it.
template <int STRIDE, int HALVES = 2>
__global__ void __launch_bounds__(HALVES * (STRIDE / 128) * 32, 1) keyswitch_tile_kernel(KeyswitchArgs A, unsigned long long count) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t row_bytes = STRIDE * 4;
    constexpr uint32_t stage_bytes = kKsT * 4 * row_bytes;          // 4 slots per digit position
    constexpr int cw = STRIDE / 128;                                 // warps across the columns (4 for n = 500, 5 for n = 630)
    constexpr int nwarps = HALVES * cw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * (size_t)stage_bytes);
    uint64_t* empty = full + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cwi = warp % cw, gh = warp / cw;                       // column block, half of the tile's ciphertexts
    const int quad = cwi * 32 + lane;
    for (int x = threadIdx.x; x < 2 * kKsT * (int)(row_bytes / 16); x += blockDim.x) {   // the zero slots of both stages
        const int slot = x / (int)(row_bytes / 16), q = x % (int)(row_bytes / 16);
        reinterpret_cast<uint4*>(smem_raw + (size_t)(slot / kKsT) * stage_bytes + (size_t)(slot % kKsT) * 4 * row_bytes)[q] = make_uint4(0, 0, 0, 0);
    }
    if (threadIdx.x == 0) {
        mbar_init(full, 1); mbar_init(full + 1, 1);
        mbar_init(empty, nwarps); mbar_init(empty + 1, nwarps);
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned long long g0 = (unsigned long long)blockIdx.x * (32 * HALVES) + gh * 32;   // first ciphertext of this warp
    const unsigned long long gl = g0 + lane;                                            // the one whose mask this lane fetches
    const bool lane_valid = gl < count;
    const int32_t* in_l = A.in + (lane_valid ? gl : 0) * A.in_stride + A.in_offset;
    constexpr uint32_t prec = 1u << (32 - (1 + kKsBasebit * kKsT));                      // keyswitch.jl:58
    const char* table = reinterpret_cast<const char*>(A.ksk);
    auto issue = [&](int i) {
        const int s = i & 1;
        mbar_arrive_expect_tx(full + s, kKsT * 3 * row_bytes);
#pragma unroll
        for (int j = 0; j < kKsT; j++)
            bulk_copy_g2s(smem_raw + (size_t)s * stage_bytes + (size_t)(j * 4 + 1) * row_bytes,
                          table + ((size_t)i * kKsT + j) * 3 * row_bytes, 3 * row_bytes, full + s);
    };
    if (threadIdx.x == 0) { issue(0); if (A.Nk > 1) issue(1); }

    uint32_t acc[32][4];
#pragma unroll
    for (int g = 0; g < 32; g++) { acc[g][0] = 0; acc[g][1] = 0; acc[g][2] = 0; acc[g][3] = 0; }

#pragma unroll 1
    for (int i = 0; i < A.Nk; i++) {
        const int s = i & 1;
        const uint32_t abar = lane_valid ? (uint32_t)__ldg(in_l + i) + prec : 0u;        // invalid ciphertext: all digits 0
        mbar_wait(full + s, (uint32_t)(i >> 1) & 1u);
        const unsigned char* rows = smem_raw + (size_t)s * stage_bytes + (size_t)quad * 16;
#pragma unroll
        for (int g = 0; g < 32; g++) {
            const uint32_t ai = __shfl_sync(0xffffffffu, abar, g);
#pragma unroll
            for (int j = 0; j < kKsT; j++) {
                const uint32_t d = (ai >> (32 - (j + 1) * kKsBasebit)) & 3u;             // keyswitch.jl:63-67
                const uint4 v = *reinterpret_cast<const uint4*>(rows + (size_t)j * 4 * row_bytes + d * row_bytes);
                acc[g][0] -= v.x; acc[g][1] -= v.y; acc[g][2] -= v.z; acc[g][3] -= v.w;   // keyswitch.jl:71-77
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        if (threadIdx.x == 0 && i + 2 < A.Nk) {
            mbar_wait(empty + s, (uint32_t)(i >> 1) & 1u);   // every warp is done with stage s
            issue(i + 2);
        }
    }
    const int c0 = quad * 4;
#pragma unroll
    for (int g = 0; g < 32; g++) {
        const unsigned long long gg = g0 + g;
        if (gg < count) {
            int32_t* o = A.out + gg * A.out_stride + A.out_offset;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int c = c0 + e;
                if (c < A.n) o[c] = (int32_t)acc[g][e];
                else if (c == A.n) {
                    int32_t* ob = A.out_b + gg * A.out_stride + A.b_offset;
                    if (A.b_mode == 0) *ob = (int32_t)((uint32_t)A.in[gg * A.in_stride + A.in_b_offset] + acc[g][e]);
                    else atomicAdd(reinterpret_cast<unsigned int*>(ob), acc[g][e]);
                }
            }
        }
    }
}

struct BlindRotateArgs {
    const double2* bk_fft;   // [n][L][2][2][NP][512]
    const double2* E;        // twiddle table, 2048 entries
    // MODE 0 (bootstrap_wo_keyswitch with fused gate prologue): lin = ka*x + kb*y + (0, cb)
    const int32_t* x; const int32_t* y;
    int32_t ka, kb, cb, mu;
    // gate_mux (gates.jl:163-171) runs its two bootstraps in one launch: gates g >= half take their inputs from
    // (x2, y2)[g - half] and the second set of prologue constants (half == 0: unused)
    const int32_t* x2; const int32_t* y2;
    int32_t ka2, kb2, cb2;
    unsigned long long half;
    // MODE 1 (raw blind_rotate on given accumulators)
    const int32_t* acc_in; const int32_t* bara_in;
    int32_t* out;            // MODE 0: [count][N+1] extracted LWE; MODE 1: [count][2][N]
    int n, n_iter, n_pad;
    unsigned long long count;
    // K3 gate -> CTA map: the first `split` CTAs hold G gates each, every later CTA `tail` (<= G) gates.  The launcher makes
    // the LAST wave of CTAs carry fewer gates each instead of leaving SMs idle (groups without a gate do no work, so such a
    // CTA finishes sooner): 1 024 gates = one wave of 148 x 4 + one wave of 148 x 3 instead of 148 x 4 + 108 x 4.
    unsigned split; int tail;
    unsigned long long* probe;   // development: clock64 phase probe of the OPT bit-3 kernel variants (else null)
    int l2_hint;                 // 1: key chunks are fetched with the L2 evict_last policy
};

// per-group shared memory: X1 + X2 + acc (+ bara, n_pad words)
__host__ __device__ constexpr int group_smem_bytes(int /*NP*/) { return (kSpectrum + kX2Elems) * 16 + 2 * kN * 4; }
// TM == 3 (output-stationary step): two X1 buffers per group, used alternately (SYNC == 1 transforms)
__host__ __device__ constexpr size_t br_group_bytes(int NP, int TM) {
    return (size_t)group_smem_bytes(NP) + (TM == 3 ? kSpectrum * 16 : 0);
}
__host__ __device__ constexpr size_t br_smem_bytes(int NP, int G, int STAGES, int n_pad, int TM = 0) {
    return (size_t)STAGES * kChunkBytes + 128 + (size_t)G * (br_group_bytes(NP, TM) + n_pad * 4);
}

// TMEM columns to allocate when the accumulators live in tensor memory: warps that share a lane quarter
// (warp % 4) get disjoint column ranges of 64*NP columns each; allocations are powers of two.
__host__ __device__ constexpr int br_tmem_cols_per_warp(int NP, int L, int TM) { return TM == 3 ? 2 * L * 32 : 64 * NP; }
__host__ __device__ constexpr int br_tmem_cols(int NP, int G, int L = 2, int TM = 1) {
    int need = ((2 * G + 3) / 4) * br_tmem_cols_per_warp(NP, L, TM), c = 32;
    while (c < need) c *= 2;
    return c;
}

// OPT: bit 3 = clock64 phase probe (development; TM == 3 only; writes A.probe); bit 4 = transforms in pairs (TM == 3, l = 2);
//      bit 7 = a dedicated producer warp walks the key ring and the compute warps only consume.  The CTA then has a
//      third warpgroup (384 threads are launched with 168 registers each); it hands its registers back
//      (setmaxnreg.dec 24) and the two compute warpgroups grow to 240 (setmaxnreg.inc), so every sub-partition holds
//      240 + 240 + 24 registers per lane — the full file — and no compute warp executes producer code or waits for a
//      stage to drain.
template <int L, int BGBIT, int NP, int G, int STAGES, int MODE, int TM = 0, int OPT = 0>
__global__ void __launch_bounds__(64 * G + ((OPT >> 7) & 1) * 128, 1) blind_rotate_kernel(BlindRotateArgs A) {
    constexpr int PROBE = (OPT >> 3) & 1, PWARP = (OPT >> 7) & 1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem_base;
    // TM: 0 = accumulators in registers, 1 = all in TMEM, 2 = component 1 in TMEM,
    //     3 = output-stationary step (extern_product_step_os; NP == 2): forward spectra in TMEM, full twiddle set
    constexpr bool kUseTmem = TM != 0;
    static_assert(TM != 3 || NP == 2, "the output-stationary step is the two-piece path");
    constexpr int kTmemCols = br_tmem_cols(NP, G, L, TM);
    static_assert(kTmemCols <= 512, "tensor memory: too many gates per CTA");
    constexpr size_t kGroupBytes = br_group_bytes(NP, TM);
    if (kUseTmem) {
        if ((threadIdx.x >> 5) == 0) tmem_alloc<kTmemCols>(&s_tmem_base);
        tmem_fence_before_sync();
    }
    double2* ring = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * kChunkBytes);
    uint64_t* empty = full + STAGES;
    unsigned char* groups = smem_raw + (size_t)STAGES * kChunkBytes + 128;

    // gates of this CTA (see BlindRotateArgs::split): groups beyond them do nothing and are not waited for by the ring
    const int cta_gates = blockIdx.x < A.split ? G : A.tail;
    const unsigned long long g_first = blockIdx.x < A.split ? (unsigned long long)blockIdx.x * G
                                                            : (unsigned long long)A.split * G + (unsigned long long)(blockIdx.x - A.split) * A.tail;
    const int n_valid = (int)(A.count - g_first < (unsigned long long)cta_gates ? A.count - g_first : (unsigned long long)cta_gates);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 2 * n_valid); }   // one arrival per consuming warp
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t tm = 0;
    if (kUseTmem) {
        tmem_fence_after_sync();
        const int warp = threadIdx.x >> 5;
        tm = s_tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * br_tmem_cols_per_warp(NP, L, TM));
    }
    BkFromRing<L, NP, STAGES, TM == 3 ? 1 : 0, PROBE, PWARP> bk{ring, full, empty, A.bk_fft, A.n_iter * 2 * L * NP, 0, 0u, threadIdx.x == 0};
    if (PWARP) {
        static_assert(!PWARP || G == 4, "register rebalancing assumes two full compute warpgroups");
        if ((threadIdx.x >> 5) >= 2 * G) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory");
            if (threadIdx.x == 64 * G) {
                if (A.l2_hint) bk.policy = l2_policy_evict_last();
                bk.produce_all();
            }
            return;   // the compute warps meet at named barriers only from here on
        }
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;" ::: "memory");
    }
    bk.prologue();   // the first STAGES-1 chunks are in flight while the gate prologue below runs

    // ---------------- consumers: one 64-thread group per gate ----------------
    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int bar_id = grp + 1;
    unsigned char* base = groups + (size_t)grp * (kGroupBytes + A.n_pad * 4);
    double2* X1 = reinterpret_cast<double2*>(base);                      // TM == 3: two buffers, used alternately
    double2* X2 = X1 + (TM == 3 ? 2 : 1) * kSpectrum;
    int32_t* acc = reinterpret_cast<int32_t*>(X2 + kX2Elems);
    int32_t* bara = acc + 2 * kN;
    const unsigned long long g = g_first + grp;
    const bool valid = grp < n_valid;   // a group without a gate skips the walk (the ring's empty barriers count the valid groups only)
    typename std::conditional<TM == 3, TwiddlesFull, Twiddles>::type w; w.load(A.E, t);

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
        // acc = (0, X^{-barb} * (mu, ..., mu))   (bootstrap.jl:54-56,78)
        const int s = (-barb) & 2047;
        for (int x = t; x < kN; x += 64) {
            acc[x] = 0;
            int yy = (x - s) & 2047;
            acc[kN + x] = (yy & 1024) ? (int32_t)(0u - (uint32_t)A.mu) : A.mu;
        }
    } else {
        const int32_t* ain = A.acc_in + g * (2 * kN);
        for (int x = t; x < 2 * kN; x += 64) acc[x] = ain[x];
        for (int i = t; i < A.n_iter; i += 64) bara[i] = A.bara_in[g * A.n + i];
    }
    group_sync(bar_id);

    long long pr[4] = {0, 0, 0, 0};
    long long t_start = 0;
    if (PROBE) t_start = clock64();
    // a group without a gate does not walk.  The vote makes the bound warp-uniform for the compiler: with a per-thread
    // bound it treats the whole iteration as divergent code (WARPSYNCs, no uniform registers) and the kernel loses 5 %
    const int n_walk = __all_sync(0xffffffffu, valid) ? A.n_iter : 0;
#pragma unroll 1
    for (int i = 0; i < n_walk; i++) {   // bootstrap.jl:19-23
        if constexpr (TM == 3 && ((OPT >> 4) & 1) && L == 2) extern_product_step_os_dual<BGBIT>(acc, bara[i], bk, w, X1, X2, tm, t, bar_id);
        else if constexpr (TM == 3) extern_product_step_os<L, BGBIT, 1, PROBE>(acc, bara[i], bk, w, X1, X2, tm, t, bar_id, pr);
        else if constexpr (TM != 0) extern_product_step_tmem<L, BGBIT, NP, (TM == 2 && NP == 2) ? 1 : 0>(acc, bara[i], bk, w, X1, X2, tm, t, bar_id);
        else extern_product_step<L, BGBIT, NP, true, true>(acc, bara[i], bk, w, X1, X2, t, bar_id);
    }

    if (PROBE && A.probe && blockIdx.x < 4 && (threadIdx.x & 31) == 0) {
        // per warp: total, forward phase, MAC passes, inverse + update, end barrier, key waits (all / first of pass / producer)
        unsigned long long* o = A.probe + ((size_t)blockIdx.x * 2 * G + (threadIdx.x >> 5)) * 8;
        o[0] = (unsigned long long)(clock64() - t_start);
        o[1] = pr[0]; o[2] = pr[1]; o[3] = pr[2]; o[4] = pr[3];
        o[5] = bk.w_full; o[6] = bk.w_first; o[7] = bk.w_empty;
    }
    if (!valid) {
    } else if (MODE == 0) {
        // tlwe_extract_sample (tlwe.jl:55-59): a = (p_0, -p_{N-1}, ..., -p_1), b = acc_b[0]
        int32_t* o = A.out + g * (kN + 1);
        for (int x = t; x < kN; x += 64) o[x] = x == 0 ? acc[0] : (int32_t)(0u - (uint32_t)acc[kN - x]);
        if (t == 0) o[kN] = acc[kN];
    } else {
        int32_t* o = A.out + g * (2 * kN);
        for (int x = t; x < 2 * kN; x += 64) o[x] = acc[x];
    }
    if (kUseTmem) {
        tmem_fence_before_sync();
        if (PWARP) asm volatile("bar.sync 15, %0;" ::"n"(64 * G) : "memory");   // the producer warp has left
        else __syncthreads();
        if ((threadIdx.x >> 5) == 0) tmem_dealloc<kTmemCols>(s_tmem_base);
    }
}

}  // namespace tfhe_b200

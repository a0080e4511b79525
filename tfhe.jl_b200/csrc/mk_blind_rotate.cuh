// mk_blind_rotate.cuh — K5: the persistent MK-TFHE blind-rotation kernel (mk_internals.jl:464-509) for sm_100a.
//
// Same organisation as K3 (blind_rotate.cuh): one CTA keeps G gates resident for all p*n iterations, one
// 64-thread group per gate, and the groups walk the expanded bootstrapping key in lockstep so that every key
// polynomial is pulled from L2 ONCE per CTA by a TMA bulk copy into a shared-memory ring.  That matters more
// here than for the single-key path: the expanded key is p*n*l*(2p+2) polynomials (393 MB with split
// transforms for 2 parties — larger than the 126 MB L2), and the one-gate-per-CTA kernel it replaces
// (mk_kernels.cuh) pulled 3.5 TB/s through L2 with 592 independent read streams.
//
// Per iteration (party, j) a gate consumes l*(3p+1) key polynomials in a fixed order that depends only on
// (p, party): for every input polynomial q and digit r — y[r,q], x[r,q] and, when q is another party's mask,
// y[r,party] again; for the b polynomial c1[r], c0[r].  The order table is built once in shared memory and the
// producer (thread 0) walks it incrementally.
//
// Accumulators: A (-> a'_party) stays in registers for the whole iteration, B (-> b') too when one 32-bit piece
// is transformed; the short-lived S (-> a'_q, q != party) and, with two-piece transforms, B live in TENSOR
// MEMORY (tmem.cuh): four gates' working sets then fit in 227 KB and nothing spills.
//
// Zero rotations (mk_internals.jl:478 skips them) are executed, as in K3: the digits of X^0*acc - acc are all
// zero, the products are exactly zero and the accumulator is unchanged — bit-identical to skipping.
#pragma once
#include "blind_rotate.cuh"
#include "mk_kernels.cuh"

namespace tfhe_b200 {

// ring of whole key polynomials (all NP pieces): NP * 8 KB per stage
template <int NP, int STAGES, int PW = 0> struct MkKeyRing {
    static constexpr int kElems = NP * kSpectrum;
    static constexpr uint32_t kBytes = (uint32_t)kElems * 16u;
    const double2* ring; uint64_t* full; uint64_t* empty;
    const double2* bk;         // [p*n][spolys] polynomials of kElems
    const int16_t* order;      // [p][cpi]: polynomial index inside one expanded sample, in consumption order
    int cpi, n, spolys, total;
    bool producer;
    // consumer cursor
    int stage = 0; uint32_t phase = 0;
    // producer cursor (next chunk to issue)
    int iss = 0, iss_stage = 0, iss_round = 0, iss_slot = 0, iss_it = 0, iss_j = 0, iss_party = 0;

    __device__ __forceinline__ void issue_next() {
        const double2* src = bk + ((size_t)iss_it * spolys + (size_t)order[iss_party * cpi + iss_slot]) * kElems;
        mbar_arrive_expect_tx(full + iss_stage, kBytes);
        bulk_copy_g2s(const_cast<double2*>(ring) + (size_t)iss_stage * kElems, src, kBytes, full + iss_stage);
        iss++;
        if (++iss_stage == STAGES) { iss_stage = 0; iss_round++; }
        if (++iss_slot == cpi) {
            iss_slot = 0; iss_it++;
            if (++iss_j == n) { iss_j = 0; iss_party++; }
        }
    }
    __device__ __forceinline__ void prologue() {
        if (producer && !PW)
            while (iss < STAGES - 1 && iss < total) issue_next();
    }
    // PW: the whole walk by one thread of a warp that does nothing else (see blind_rotate.cuh, BkFromRing DIST == 2)
    __device__ __forceinline__ void produce_all() {
        while (iss < total) {
            if (iss_round >= 1) mbar_wait(empty + iss_stage, (uint32_t)(iss_round - 1) & 1u);
            issue_next();
        }
    }
    uint32_t ready_next = 0;   // PW: early test of the next polynomial's barrier (an mbarrier test costs 100-150 cycles)
    __device__ __forceinline__ const double2* acquire() {
        if (PW) {
            if (!ready_next) mbar_wait_poll(full + stage, phase);
            const int ns = stage + 1 == STAGES ? 0 : stage + 1;
            uint32_t done;
            asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(full + ns)), "r"(ns == 0 ? phase ^ 1u : phase) : "memory");
            ready_next = done;
            return ring + (size_t)stage * kElems;
        }
        if (producer && iss < total) {
            // refill the stage released one chunk ago (all groups passed it an FFT ago)
            if (iss_round >= 1) mbar_wait(empty + iss_stage, (uint32_t)(iss_round - 1) & 1u);
            issue_next();
        }
        mbar_wait(full + stage, phase);
        return ring + (size_t)stage * kElems;
    }
    __device__ __forceinline__ void release() {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(empty + stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
};

template <int NP>
__device__ __forceinline__ void mac_spectrum_smem(double2 (&o)[NP][8], const double2 (&a)[8], const double2* b) {
#pragma unroll
    for (int pc = 0; pc < NP; pc++)
#pragma unroll
        for (int q = 0; q < 8; q++) cmac(o[pc][q], a[q], b[(pc * 8 + q) * 64]);
}

// o (NP spectrum slices in this thread's TMEM row at column `col`) += a * b;  `first` starts from zero instead
template <int NP>
__device__ __forceinline__ void mac_spectrum_tmem(uint32_t col, const double2 (&a)[8], const double2* b, bool first) {
    if (!first) tmem_wait_st();   // this thread's previous stores to these columns have landed
#pragma unroll
    for (int pc = 0; pc < NP; pc++) {
        double2 o[8];
        if (first) {
#pragma unroll
            for (int e = 0; e < 8; e++) o[e] = make_double2(0.0, 0.0);
        } else {
            tmem_load_spectrum(col + (uint32_t)(pc * 32), o);
        }
#pragma unroll
        for (int e = 0; e < 8; e++) cmac(o[e], a[e], b[(pc * 8 + e) * 64]);
        tmem_store_spectrum(col + (uint32_t)(pc * 32), o);
    }
}

// mk_tgsw_extern_mul (mk_internals.jl:348-391) on acc[(p+1)][N] in shared memory, key from the ring, S in TMEM
// (`tmS`: this thread's TMEM row, NP*32 columns for S and, when BT, NP*32 more for B: with two-piece transforms
// A and B together would need 128 registers and spill).
template <int L, int BGBIT, int NP, bool BT, class KEY>
__device__ __forceinline__ void mk_extern_product_step_ring(int32_t* acc, int p, int party, int abar, KEY& key,
                                                            const Twiddles& w, double2* X1, double2* X2, uint32_t tmS,
                                                            int t, int bar_id) {
    constexpr uint32_t offset = decomp_offset<L, BGBIT>();
    const int s = abar & 2047;
    const uint32_t tmB = tmS + (uint32_t)(NP * 32);
    double2 A[NP][8], B[BT ? 1 : NP][8];
#pragma unroll
    for (int pc = 0; pc < NP; pc++)
#pragma unroll
        for (int q = 0; q < 8; q++) {
            A[pc][q] = make_double2(0.0, 0.0);
            if (!BT) B[pc % (BT ? 1 : NP)][q] = make_double2(0.0, 0.0);
        }

#pragma unroll 1
    for (int q = 0; q <= p; q++) {
        int32_t* poly = acc + q * kN;
        uint32_t tl[8], th[8];
#pragma unroll
        for (int m = 0; m < 8; m++) {
            int j = t + 64 * m;
            tl[m] = (uint32_t)rot_coeff(poly, j, s) - (uint32_t)poly[j] + offset;          // mk_internals.jl:468
            th[m] = (uint32_t)rot_coeff(poly, j + 512, s) - (uint32_t)poly[j + 512] + offset;
        }
        const bool side = q < p && q != party;   // this input also feeds a'_q through y[.,party]
#pragma unroll 1
        for (int r = 0; r < L; r++) {
            double2 a[8];
#pragma unroll
            for (int m = 0; m < 8; m++)
                a[m] = make_double2(digit_f64<BGBIT>(tl[m], r), -digit_f64<BGBIT>(th[m], r));   // :356-357
            fft512_forward(a, w, X1, X2, t, bar_id);                                                    // :368-369
            mac_spectrum_smem<NP>(A, a, key.acquire() + t);   // q < p: y[r,q] (:375-376);  q == p: c1[r] (:377-378)
            key.release();
            {                                                 // q < p: x[r,q] (:384-385);  q == p: c0[r] (:386-387)
                const double2* b = key.acquire() + t;
                if constexpr (BT) mac_spectrum_tmem<NP>(tmB, a, b, q == 0 && r == 0);
                else mac_spectrum_smem<BT ? 1 : NP>(B, a, b);
                key.release();
            }
            if (side) {                                       // y[r,party] (:379-380)
                mac_spectrum_tmem<NP>(tmS, a, key.acquire() + t, r == 0);
                key.release();
            }
        }
        if (side) {
            double2 o[NP][8];
            tmem_wait_st();
#pragma unroll
            for (int pc = 0; pc < NP; pc++) tmem_load_spectrum(tmS + (uint32_t)(pc * 32), o[pc]);
            group_sync(bar_id);   // all reads of acc[q] and of X2 (last forward) are done
            finish_poly<NP, true>(o, poly, w, X1, X2, t, bar_id);
            group_sync(bar_id);   // X1 free again before the next forward transform
        }
    }
    group_sync(bar_id);
    finish_poly<NP, true>(A, acc + party * kN, w, X1, X2, t, bar_id);
    if constexpr (BT) {
        tmem_wait_st();
#pragma unroll
        for (int pc = 0; pc < NP; pc++) tmem_load_spectrum(tmB + (uint32_t)(pc * 32), A[pc]);
        finish_poly<NP, true>(A, acc + p * kN, w, X1, X2, t, bar_id);
    } else {
        finish_poly<BT ? 1 : NP, true>(B, acc + p * kN, w, X1, X2, t, bar_id);
    }
    group_sync(bar_id);
}

__host__ __device__ inline int mk_chunks_per_iter(int L, int p) { return L * (3 * p + 1); }
__host__ __device__ inline size_t mk_order_bytes(int L, int p) { return ((size_t)p * mk_chunks_per_iter(L, p) * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t mk_group_bytes(int p, int n) {
    return (size_t)(kSpectrum + kX2Elems) * 16 + (size_t)(p + 1) * kN * 4 + (((size_t)p * n * 2 + 15) & ~(size_t)15);
}
__host__ __device__ inline size_t mk_ring_smem_bytes(int L, int p, int n, int NP, int G, int STAGES) {
    return (size_t)STAGES * NP * kSpectrum * 16 + 128 + mk_order_bytes(L, p) + (size_t)G * mk_group_bytes(p, n);
}
__host__ __device__ constexpr int mk_tmem_cols_per_warp(int NP) { return NP == 2 ? 128 : 32; }   // S (+ B when NP == 2)
__host__ __device__ constexpr int mk_tmem_cols(int NP, int G) {
    int need = ((2 * G + 3) / 4) * mk_tmem_cols_per_warp(NP), c = 32;
    while (c < need) c *= 2;
    return c;
}

// PW: a dedicated producer warpgroup walks the key ring (see blind_rotate.cuh, OPT bit 7).  With G = 4 the CTA is launched
// with 384 threads at 168 registers and rebalanced to 240 / 24 by setmaxnreg; with G = 2 (256 threads, 255 registers) the
// sub-partitions hold one compute and one producer warp each and nothing needs rebalancing.
template <int L, int BGBIT, int NP, int G, int STAGES, int PW = 0>
__global__ void __launch_bounds__(64 * G + PW * 128, 1) mk_blind_rotate_ring_kernel(MKBlindRotateArgs M) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_tmem_base;
    using Ring = MkKeyRing<NP, STAGES, PW>;
    const int p = M.p, n = M.n;
    const int cpi = mk_chunks_per_iter(L, p);
    if ((threadIdx.x >> 5) == 0) tmem_alloc<mk_tmem_cols(NP, G)>(&s_tmem_base);
    tmem_fence_before_sync();

    double2* ring = reinterpret_cast<double2*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * Ring::kBytes);
    uint64_t* empty = full + STAGES;
    int16_t* order = reinterpret_cast<int16_t*>(smem_raw + (size_t)STAGES * Ring::kBytes + 128);
    unsigned char* groups = reinterpret_cast<unsigned char*>(order) + mk_order_bytes(L, p);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 2 * G); }
        mbar_fence_init();
    }
    // consumption order of the key polynomials of one iteration, per party
    for (int e = threadIdx.x; e < p * cpi; e += blockDim.x) {
        const int party = e / cpi, slot = e - party * cpi;
        int idx = -1, pos = 0;
        for (int q = 0; q <= p && idx < 0; q++) {
            const int per_r = (q < p && q != party) ? 3 : 2;
            if (slot < pos + per_r * L) {
                const int r = (slot - pos) / per_r, which = (slot - pos) - r * per_r;
                if (q < p) idx = which == 0 ? mk_yi(L, p, r, q) : which == 1 ? mk_xi(L, p, r, q) : mk_yi(L, p, r, party);
                else idx = which == 0 ? mk_c1i(L, p, r) : mk_c0i(L, p, r);
            }
            pos += per_r * L;
        }
        order[e] = (int16_t)idx;
    }
    __syncthreads();
    tmem_fence_after_sync();
    const int warp = threadIdx.x >> 5;
    const uint32_t tmS = s_tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * mk_tmem_cols_per_warp(NP));

    Ring key{ring, full, empty, M.bk_fft, order, cpi, n, L * (2 * p + 2), p * n * cpi, threadIdx.x == 0};
    if (PW) {
        if ((threadIdx.x >> 5) >= 2 * G) {
            if (G == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;" ::: "memory");
            if (threadIdx.x == 64 * G) key.produce_all();
            return;   // the compute warps meet at named barriers only from here on
        }
        if (G == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 240;" ::: "memory");
    }
    key.prologue();

    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int bar_id = grp + 1;
    unsigned char* base = groups + (size_t)grp * mk_group_bytes(p, n);
    double2* X1 = reinterpret_cast<double2*>(base);
    double2* X2 = X1 + kSpectrum;
    int32_t* acc = reinterpret_cast<int32_t*>(X2 + kX2Elems);
    int16_t* bara = reinterpret_cast<int16_t*>(acc + (p + 1) * kN);
    const unsigned long long g = (unsigned long long)blockIdx.x * G + grp;
    const bool valid = g < M.count;   // a group without a gate still walks the ring (on zeros) to keep the lockstep
    Twiddles w; w.load(M.E, t);

    if (!valid) {
        for (int x = t; x < (p + 1) * kN; x += 64) acc[x] = 0;
        for (int i = t; i < p * n; i += 64) bara[i] = 0;
    } else {
        const size_t wct = (size_t)p * n + 1;
        const int32_t* xr = M.x + g * wct;
        const int32_t* yr = M.y ? M.y + g * wct : nullptr;
        for (int i = t; i < p * n; i += 64) {                                       // mk_gates.jl:8-10, mk_internals.jl:503
            uint32_t v = (uint32_t)M.ka * (uint32_t)xr[i];
            if (yr) v += (uint32_t)M.kb * (uint32_t)yr[i];
            bara[i] = (int16_t)modswitch2048((int32_t)v);
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
    }
    group_sync(bar_id);

#pragma unroll 1
    for (int party = 0; party < p; party++)                                         // :475
#pragma unroll 1
        for (int j = 0; j < n; j++)                                                 // :476
            mk_extern_product_step_ring<L, BGBIT, NP, NP == 2>(acc, p, party, (int)bara[party * n + j], key, w, X1, X2, tmS, t, bar_id);

    if (valid) {   // mk_tlwe_extract_sample (mk_internals.jl:88-95)
        int32_t* o = M.out + g * ((size_t)p * kN + 1);
        for (int q = 0; q < p; q++)
            for (int x = t; x < kN; x += 64)
                o[q * kN + x] = x == 0 ? acc[q * kN] : (int32_t)(0u - (uint32_t)acc[q * kN + kN - x]);
        if (t == 0) o[p * kN] = acc[p * kN];
    }
    tmem_fence_before_sync();
    if (PW) asm volatile("bar.sync 15, %0;" ::"n"(64 * G) : "memory");   // the producer warpgroup has left
    else __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc<mk_tmem_cols(NP, G)>(s_tmem_base);
}

}  // namespace tfhe_b200

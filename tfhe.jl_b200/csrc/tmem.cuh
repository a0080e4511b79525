// tmem.cuh — Blackwell tensor memory (TMEM, 256 KB per SM) used as per-thread scratch for the FP64 spectrum
// accumulators of the external product.  No tensor-core instruction is involved: tcgen05.st / tcgen05.ld with
// the 32x32b shape give every thread of a warp a private TMEM row (lane = 32*(warp%4) + laneid) addressed by
// column, and that traffic runs on its own datapath (SASS STTM / LDTM), not on the LSU / shared-memory pipe
// that the FFT exchanges saturate.  Measured on B200 (tools/tmem_test.cu): read-modify-write of 64 columns per
// thread sustains ~155 B/clk/SM in each direction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tfhe_b200 {

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 4 complex doubles (16 columns) of this thread's row
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const double2* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr),
                   "r"(__double2loint(v[0].x)), "r"(__double2hiint(v[0].x)), "r"(__double2loint(v[0].y)), "r"(__double2hiint(v[0].y)),
                   "r"(__double2loint(v[1].x)), "r"(__double2hiint(v[1].x)), "r"(__double2loint(v[1].y)), "r"(__double2hiint(v[1].y)),
                   "r"(__double2loint(v[2].x)), "r"(__double2hiint(v[2].x)), "r"(__double2loint(v[2].y)), "r"(__double2hiint(v[2].y)),
                   "r"(__double2loint(v[3].x)), "r"(__double2hiint(v[3].x)), "r"(__double2loint(v[3].y)), "r"(__double2hiint(v[3].y))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4_raw(uint32_t taddr, int (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
// one spectrum slice = 8 complex doubles = 32 columns
__device__ __forceinline__ void tmem_store_spectrum(uint32_t taddr, const double2 (&v)[8]) {
    tmem_st4(taddr, &v[0]);
    tmem_st4(taddr + 16, &v[4]);
}
__device__ __forceinline__ void tmem_load_spectrum(uint32_t taddr, double2 (&v)[8]) {
    int r0[16], r1[16];
    tmem_ld4_raw(taddr, r0);
    tmem_ld4_raw(taddr + 16, r1);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[i] = make_double2(__hiloint2double(r0[4 * i + 1], r0[4 * i]), __hiloint2double(r0[4 * i + 3], r0[4 * i + 2]));
        v[4 + i] = make_double2(__hiloint2double(r1[4 * i + 1], r1[4 * i]), __hiloint2double(r1[4 * i + 3], r1[4 * i + 2]));
    }
}

}  // namespace tfhe_b200

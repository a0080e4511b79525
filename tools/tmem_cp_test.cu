// tmem_cp_test.cu — development probe: (1) does tcgen05.cp.64x128b.warpx2::02_13 with a no-swizzle descriptor put row t
// of a [e][64 rows][16 B] shared-memory block into TMEM lane t AND lane t+64, columns 4e..4e+3?  (2) how fast are
// tcgen05.ld (32x32b), tcgen05.cp and LDS.128 alone and together on one SM?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_cp_test tools/tmem_cp_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done; const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t t) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_cp_64x128b_02_13(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void tmem_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_u32(smem) & 0x3ffff) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

// ---- (1) layout check -------------------------------------------------------------------------------
// out[thread][64]: the 64 columns every thread reads back from its own TMEM lane
__global__ void __launch_bounds__(256, 1) k_layout(uint32_t* out, uint32_t lbo, uint32_t sbo) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_base;
    __shared__ __align__(8) uint64_t bar;
    uint32_t* words = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 4096; i += 256) words[i] = (uint32_t)i;   // word index = ((e*64 + row)*4 + j)
    if ((threadIdx.x >> 5) == 0) tmem_alloc<64>(&s_base);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (tcgen05.cp)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base;
    if (threadIdx.x == 0) {
        for (int e = 0; e < 16; e++) tmem_cp_64x128b_02_13(base + 4 * e, make_desc(smem + e * 1024, lbo, sbo));
        tmem_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = threadIdx.x >> 5;
    const uint32_t tm = base + ((uint32_t)(32 * (warp & 3)) << 16);
    uint32_t r[16];
    for (int c = 0; c < 4; c++) {
        ld16(tm + 16 * c, r); wait_ld();
        for (int i = 0; i < 16; i++) out[threadIdx.x * 64 + c * 16 + i] = r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc<64>(base);
}

// ---- (2) throughput ---------------------------------------------------------------------------------
// mode bit 0: every warp streams tcgen05.ld (64 columns per step); bit 1: every warp streams LDS.128 (16 per step);
// bit 2: thread 0 streams tcgen05.cp (16 KB per step).  Returns cycles of the slowest warp via out.
__global__ void __launch_bounds__(256, 1) k_rate(unsigned long long* out, uint32_t* sink, int mode, int steps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_base;
    __shared__ __align__(8) uint64_t bar;
    uint4* q = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 4096; i += 256) q[i] = make_uint4(i, i + 1, i + 2, i + 3);   // 64 KB
    if ((threadIdx.x >> 5) == 0) tmem_alloc<512>(&s_base);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base;
    const int warp = threadIdx.x >> 5;
    const uint32_t tm = base + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 256;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int s = 0; s < steps; s++) {
        if ((mode & 4) && threadIdx.x == 0) {
            const uint32_t col = 64 * (s & 3) + 256 * 0;   // overwrites columns nobody reads in this mode mix: fine for a rate probe
            for (int e = 0; e < 16; e++) tmem_cp_64x128b_02_13(base + col + 4 * e, make_desc(smem + (s & 3) * 16384 + e * 1024, 128, 128));
        }
        if (mode & 1) {
            uint32_t r0[16], r1[16], r2[16], r3[16];
            const uint32_t c = tm + 64 * (s & 3);
            ld16(c, r0); ld16(c + 16, r1); ld16(c + 32, r2); ld16(c + 48, r3);
            wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i++) acc ^= r0[i] ^ r1[i] ^ r2[i] ^ r3[i];
        }
        if (mode & 2) {
            const uint4* p = q + (threadIdx.x & 63) + ((s & 3) * 1024);
#pragma unroll
            for (int e = 0; e < 16; e++) { uint4 v = p[e * 64]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        }
    }
    if ((mode & 4) && threadIdx.x == 0) { tmem_commit(&bar); mbar_wait(&bar, 0); }
    const long long t1 = clock64();
    sink[blockIdx.x * 256 + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) tmem_dealloc<512>(base);
}

int main() {
    uint32_t* d_out; cudaMalloc(&d_out, 256 * 64 * 4);
    std::vector<uint32_t> h(256 * 64);
    printf("{\"layout\": [");
    const uint32_t cand[][2] = {{128, 128}, {16, 128}, {0, 128}};
    bool first = true;
    for (auto& c : cand) {
        cudaMemset(d_out, 0xff, 256 * 64 * 4);
        k_layout<<<1, 256, 16384>>>(d_out, c[0], c[1]);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s{\"lbo\": %u, \"sbo\": %u, \"error\": \"%s\"}", first ? "" : ", ", c[0], c[1], cudaGetErrorString(e)); first = false; break; }
        cudaMemcpy(h.data(), d_out, 256 * 64 * 4, cudaMemcpyDeviceToHost);
        // expectation: thread (warp w, lane l) owns TMEM lane 32*(w%4)+l; row = lane % 64; column 4e+j = word ((e*64+row)*4+j)
        int bad = 0;
        for (int t = 0; t < 256; t++) {
            const int lane = 32 * ((t >> 5) & 3) + (t & 31), row = lane % 64;
            for (int col = 0; col < 64; col++) bad += h[t * 64 + col] != (uint32_t)(((col / 4) * 64 + row) * 4 + col % 4);
        }
        printf("%s{\"lbo\": %u, \"sbo\": %u, \"mismatches\": %d, \"t0\": [%u,%u,%u,%u,%u,%u,%u,%u], \"t1\": [%u,%u,%u,%u], \"t33\": [%u,%u,%u,%u], \"t70\": [%u,%u,%u,%u]}",
               first ? "" : ", ", c[0], c[1], bad, h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[64], h[65], h[66], h[67],
               h[33 * 64], h[33 * 64 + 1], h[33 * 64 + 2], h[33 * 64 + 3], h[70 * 64], h[70 * 64 + 1], h[70 * 64 + 2], h[70 * 64 + 3]);
        first = false;
    }
    printf("]");
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, steps = 2000;
    unsigned long long* d_cyc; cudaMalloc(&d_cyc, sms * 8);
    uint32_t* d_sink; cudaMalloc(&d_sink, sms * 256 * 4);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    printf(", \"rates\": [");
    for (int mode = 1; mode < 8; mode++) {
        k_rate<<<sms, 256, 65536>>>(d_cyc, d_sink, mode, steps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"mode\": %d, \"error\": \"%s\"}", mode, cudaGetErrorString(e)); break; }
        std::vector<unsigned long long> cyc(sms);
        cudaMemcpy(cyc.data(), d_cyc, sms * 8, cudaMemcpyDeviceToHost);
        double mean = 0; for (auto v : cyc) mean += (double)v; mean /= sms;
        // bytes per step per SM: ld = 256 threads * 64 cols * 4 B = 64 KB; lds = 256 * 16 * 16 = 64 KB; cp = 16 KB (smem side)
        printf("%s{\"mode\": %d, \"cycles_per_step\": %.1f, \"ldtm_B_per_clk\": %.1f, \"lds_B_per_clk\": %.1f, \"cp_B_per_clk\": %.1f}", mode > 1 ? ", " : "",
               mode, mean / steps, (mode & 1) ? 65536.0 * steps / mean : 0.0, (mode & 2) ? 65536.0 * steps / mean : 0.0,
               (mode & 4) ? 16384.0 * steps / mean : 0.0);
    }
    printf("]}\n");
    return 0;
}

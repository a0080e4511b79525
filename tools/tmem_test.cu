// tmem_test.cu — can tensor memory (TMEM) serve as per-thread scratch for FP64 accumulators?
// Checks tcgen05.alloc / st / ld / dealloc with the 32x32b shape (one private TMEM row per thread) and
// measures st+ld round-trip throughput.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int COLS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// 16 x 32-bit columns of this thread's TMEM row
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kCols = 256;

// every thread: write a pattern to its 64 columns, read back, accumulate in a loop; report mismatches
__global__ void __launch_bounds__(512, 1) tmem_kernel(int iters, unsigned* errors, unsigned long long* cycles, double* sink) {
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<kCols>(&s_base);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base;
    // this thread's row: lane quarter of the warp, column range by warp / 4
    const uint32_t taddr = base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 64);
    uint32_t v[16];
    unsigned bad = 0;
    // correctness: 4 blocks of 16 columns
    for (int b = 0; b < 4; b++) {
        for (int i = 0; i < 16; i++) v[i] = 0x9e3779b9u * (threadIdx.x * 64 + b * 16 + i + 1) + blockIdx.x;
        tmem_st16(taddr + b * 16, v);
    }
    tmem_wait_st();
    for (int b = 0; b < 4; b++) {
        tmem_ld16(taddr + b * 16, v);
        tmem_wait_ld();
        for (int i = 0; i < 16; i++) bad += v[i] != 0x9e3779b9u * (threadIdx.x * 64 + b * 16 + i + 1) + blockIdx.x;
    }
    if (bad) atomicAdd(errors, bad);
    // throughput: read-modify-write of 64 columns per iteration (the accumulator update pattern)
    double acc = 0;
    __syncthreads();
    unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            tmem_ld16(taddr + b * 16, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                double d = __hiloint2double((int)v[i + 1], (int)v[i]);
                d = fma(d, 1.0000001, 1e-9);
                v[i] = (uint32_t)__double2loint(d); v[i + 1] = (uint32_t)__double2hiint(d);
                acc += d;
            }
            tmem_st16(taddr + b * 16, v);
        }
        tmem_wait_st();
    }
    unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc<kCols>(base);
}

int main() {
    unsigned* d_err; unsigned long long* d_cyc; double* d_sink;
    const int blocks = 148, threads = 512, iters = 2000;
    cudaMalloc(&d_err, 4); cudaMemset(d_err, 0, 4);
    cudaMalloc(&d_cyc, blocks * 8); cudaMalloc(&d_sink, (size_t)blocks * threads * 8);
    tmem_kernel<<<blocks, threads>>>(iters, d_err, d_cyc, d_sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned err = 0; unsigned long long cyc[148];
    cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost); cudaMemcpy(cyc, d_cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double bytes = (double)threads * 64 * 4 * iters;   // per direction per SM
    printf("{\"cuda\": \"%s\", \"mismatches\": %u, \"cycles\": %llu, \"tmem_rmw_bytes_per_clk_per_sm_each_way\": %.1f}\n",
           cudaGetErrorString(e), err, cyc[0], bytes / (double)cyc[0]);
    return e != cudaSuccess || err != 0;
}

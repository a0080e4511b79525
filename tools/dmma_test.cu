// dmma_test.cu — is the FP64 tensor instruction (mma.sync.m8n8k4.f64, SASS DMMA) a second FP64 pipe on B200?
// Times (A) 8 DFMA chains per thread, (D) 4 independent DMMA chains per warp, (B) both interleaved.
// B ~ max(A, D): the two run side by side and a matrix-shaped part of the blind rotation (the per-frequency 4x4 complex
// multiply-accumulate of tgsw_extern_mul) could move off the DFMA pipe.  B ~ A + D: one pipe, nothing to gain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_test tools/dmma_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ITERS 4096
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE> __global__ void k(double* out, double a, double b) {
    double x[8], c[4][2];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; i++) { c[i][0] = threadIdx.x * 1e-4; c[i][1] = i; }
    const double fa = 1e-3 * (threadIdx.x & 3), fb = 1e-3 * (threadIdx.x >> 2 & 7);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE != 2) x[i] = fma(x[i], a, b);
            if (MODE != 0 && (i & 1) == 0) dmma(c[i >> 1][0], c[i >> 1][1], fa, fb);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(int threads, double* o) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, threads>>>(o, 1.0000001, 1e-9); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<148, threads>>>(o, 1.0000001, 1e-9); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double* o; cudaMalloc(&o, 148 * 1024 * 8);
    for (int threads : {128, 256, 512, 1024}) {
        float A = run<0>(threads, o), B = run<1>(threads, o), D = run<2>(threads, o);
        // per SM and clock (1.965 GHz): DFMA lanes = threads * ITERS * 8; DMMA = (threads/32) * ITERS * 4 instructions of 256 FMA
        const double clk = 1.965e6;   // cycles per ms
        printf("{\"threads_per_sm\": %d, \"dfma_only_ms\": %.3f, \"dfma_plus_dmma_ms\": %.3f, \"dmma_only_ms\": %.3f, "
               "\"dfma_fma_per_clk_sm\": %.1f, \"dmma_fma_per_clk_sm\": %.1f}\n",
               threads, A, B, D, threads * (double)ITERS * 8 / (A * clk), (threads / 32) * (double)ITERS * 4 * 256 / (D * clk));
    }
    return 0;
}

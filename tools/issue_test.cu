// issue_test.cu — does an FP64 instruction block the warp scheduler's issue port for both of its cycles?
// Times (A) 8 DFMA chains, (C) 8 integer chains (LOP3/IADD), (B) both interleaved.  B ~ max(A, C): integer work
// hides in the FP64 shadow.  B ~ A + C: every non-FP64 instruction costs an issue slot on top.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define ITERS 4096
template <int MODE> __global__ void k(double* out, uint32_t* iout, double a, double b, uint32_t m) {
    double x[8]; uint32_t y[8];
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE != 2) x[i] = fma(x[i], a, b);
            if (MODE != 0) y[i] = (y[i] ^ m) + (y[i] >> 3);
        }
    }
    double s = 0; uint32_t t = 0;
    for (int i = 0; i < 8; i++) { s += x[i]; t += y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> float run(int threads, double* o, uint32_t* io) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, threads>>>(o, io, 1.0000001, 1e-9, 0x5bd1e995u); cudaDeviceSynchronize();
    cudaEventRecord(a); k<MODE><<<148, threads>>>(o, io, 1.0000001, 1e-9, 0x5bd1e995u); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double* o; uint32_t* io; cudaMalloc(&o, 148 * 1024 * 8); cudaMalloc(&io, 148 * 1024 * 4);
    for (int threads : {128, 256, 512}) {
        float A = run<0>(threads, o, io), B = run<1>(threads, o, io), C = run<2>(threads, o, io);
        printf("{\"threads_per_sm\": %d, \"dfma_only_ms\": %.3f, \"dfma_plus_int_ms\": %.3f, \"int_only_ms\": %.3f}\n", threads, A, B, C);
    }
    return 0;
}

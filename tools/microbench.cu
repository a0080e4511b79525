// microbench.cu — measures the B200 pipe rates the blind-rotation roofline depends on:
// FP64 FMA/ADD, INT32 IMAD / IMAD.WIDE, shared-memory LDS.128 bandwidth, SHFL rate.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench tools/microbench.cu
// Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
template <int ILP> __global__ void k_dfma(double* out, double a, double b) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_dadd(double* out, double a) {
    double x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = x[i] + a;
    double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = x[i] * a + b;
    uint32_t s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_imadwide(uint64_t* out, uint32_t a) {
    uint64_t x[ILP];
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = (uint64_t)(uint32_t)x[i] * a + x[i];
    uint64_t s = 0; for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lds128(double* out) {
    __shared__ double2 buf[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) { double2 v = buf[(idx + u * 64) & 1023]; acc.x += v.x; acc.y += v.y; }
        idx = (idx + 1) & 1023;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}
__global__ void k_shfl(uint32_t* out) {
    uint32_t x = threadIdx.x, s = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int u = 1; u <= 8; u++) { x = __shfl_xor_sync(0xffffffffu, x, u); s += x; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_f2i(long long* out, double a) {
    double x[8]; long long s = 0;
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1.5 + i;
    for (int it = 0; it < ITERS; it++)
#pragma unroll
        for (int i = 0; i < 8; i++) { s += __double2ll_rn(x[i]); x[i] += a; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, blocks = sms * 4, threads = 512;
    void* buf; cudaMalloc(&buf, (size_t)blocks * threads * 8);
    double n = (double)blocks * threads * ITERS;
    float t;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, p.clockRate);
    t = timeit([&] { k_dfma<8><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9); });
    printf(", \"dfma_tflops\": %.2f, \"dfma_per_clk_per_sm_at_max\": %.1f", n * 8 * 2 / t / 1e9, n * 8 / (t * 1e-3) / sms / (p.clockRate * 1e3));
    t = timeit([&] { k_dadd<8><<<blocks, threads>>>((double*)buf, 1e-9); });
    printf(", \"dadd_gops\": %.1f", n * 8 / t / 1e6);
    t = timeit([&] { k_imad<8><<<blocks, threads>>>((uint32_t*)buf, 3, 7); });
    printf(", \"imad_gops\": %.1f", n * 8 / t / 1e6);
    t = timeit([&] { k_imadwide<8><<<blocks, threads>>>((uint64_t*)buf, 12345); });
    printf(", \"imadwide_gops\": %.1f", n * 8 / t / 1e6);
    t = timeit([&] { k_lds128<<<blocks, threads>>>((double*)buf); });
    printf(", \"lds128_tbps\": %.2f, \"lds_bytes_per_clk_per_sm_at_max\": %.1f", n * 8 * 16 / t / 1e9, n * 8 * 16 / (t * 1e-3) / sms / (p.clockRate * 1e3));
    t = timeit([&] { k_shfl<<<blocks, threads>>>((uint32_t*)buf); });
    printf(", \"shfl_gops\": %.1f", n * 8 / t / 1e6);
    t = timeit([&] { k_f2i<<<blocks, threads>>>((long long*)buf, 0.25); });
    printf(", \"f2i64_gops\": %.1f", n * 8 / t / 1e6);
    printf("}\n");
    return 0;
}

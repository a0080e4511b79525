// dsmem_test.cu — development probe for a two-CTA latency kernel: how long does it take a CTA pair to swap 16 KB
// (two 512-point complex-double spectra each way) through distributed shared memory, every iteration?
//   mode 0  st.async (16 B per store, completes bytes on the peer's mbarrier)
//   mode 1  cp.async.bulk shared::cta -> shared::cluster, 8 KB per copy, completes on the peer's mbarrier
//   mode 2  plain st.shared::cluster + barrier.cluster (release/acquire)
//   mode 3  no copy: barrier.cluster alone (cost of the barrier)
//   mode 4  ld.shared::cluster by the consumer after a barrier.cluster
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/dsmem_test tools/dsmem_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done; const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_async16(uint32_t raddr, double2 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "l"(__double_as_longlong(v.x)), "l"(__double_as_longlong(v.y)), "r"(rbar) : "memory");
}
__device__ __forceinline__ void bulk_s2s(uint32_t rdst, uint32_t lsrc, uint32_t bytes, uint32_t rbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rdst), "r"(lsrc), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void st_cluster16(uint32_t raddr, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(raddr), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double2 ld_cluster16(uint32_t raddr) {
    double2 v; asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(raddr) : "memory"); return v;
}

constexpr int kSpec = 512;   // double2 per spectrum
// smem: src[2][512] (own spectra), land[2 buffers][2][512], mbar[2]
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_swap(int mode, int iters, long long* cycles, double* check) {
    extern __shared__ __align__(128) unsigned char smem[];
    double2* src = reinterpret_cast<double2*>(smem);
    double2* land = src + 2 * kSpec;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(land + 4 * kSpec);
    const uint32_t rank = cta_rank(), peer = rank ^ 1;
    const int t = threadIdx.x & 63, grp = threadIdx.x >> 6;
    for (int i = threadIdx.x; i < 2 * kSpec; i += 128) src[i] = make_double2(rank * 10000.0 + i, -(double)i);
    for (int i = threadIdx.x; i < 4 * kSpec; i += 128) land[i] = make_double2(0.0, 0.0);
    if (threadIdx.x < 2) mbar_init(mbar + threadIdx.x, 1);
    if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    cluster_sync();
    const uint32_t r_land = mapa(smem_u32(land), peer), r_bar = mapa(smem_u32(mbar), peer), r_src = mapa(smem_u32(src), peer);
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int buf = it & 1;
        const uint32_t par = (uint32_t)(it >> 1) & 1u;
        if (mode == 0) {
            if (threadIdx.x == 0) mbar_expect(mbar + buf, 2 * kSpec * 16);
#pragma unroll
            for (int e = 0; e < 8; e++)
                st_async16(r_land + ((buf * 2 + grp) * kSpec + e * 64 + t) * 16, src[grp * kSpec + e * 64 + t], r_bar + buf * 8);
            mbar_wait(mbar + buf, par);
        } else if (mode == 1) {
            if (threadIdx.x == 0) mbar_expect(mbar + buf, 2 * kSpec * 16);
            if (t == 0) bulk_s2s(r_land + (buf * 2 + grp) * kSpec * 16, smem_u32(src + grp * kSpec), kSpec * 16, r_bar + buf * 8);
            mbar_wait(mbar + buf, par);
        } else if (mode == 5) {   // group 0 by st.async, group 1 by one bulk copy: do the two paths add up?
            if (threadIdx.x == 0) mbar_expect(mbar + buf, 2 * kSpec * 16);
            if (grp == 0) {
#pragma unroll
                for (int e = 0; e < 8; e++)
                    st_async16(r_land + ((buf * 2 + grp) * kSpec + e * 64 + t) * 16, src[grp * kSpec + e * 64 + t], r_bar + buf * 8);
            } else if (t == 0) bulk_s2s(r_land + (buf * 2 + grp) * kSpec * 16, smem_u32(src + grp * kSpec), kSpec * 16, r_bar + buf * 8);
            mbar_wait(mbar + buf, par);
        } else if (mode == 6) {   // half the volume (one spectrum each way) by st.async: is the cost linear in bytes?
            if (threadIdx.x == 0) mbar_expect(mbar + buf, kSpec * 16);
            if (grp == 0) {
#pragma unroll
                for (int e = 0; e < 8; e++)
                    st_async16(r_land + ((buf * 2 + grp) * kSpec + e * 64 + t) * 16, src[grp * kSpec + e * 64 + t], r_bar + buf * 8);
            }
            mbar_wait(mbar + buf, par);
        } else if (mode == 2) {
#pragma unroll
            for (int e = 0; e < 8; e++)
                st_cluster16(r_land + ((buf * 2 + grp) * kSpec + e * 64 + t) * 16, src[grp * kSpec + e * 64 + t]);
            cluster_sync();
        } else if (mode == 3) {
            cluster_sync();
        } else {
            cluster_sync();
#pragma unroll
            for (int e = 0; e < 8; e++)
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    double2 v = ld_cluster16(r_src + (q * kSpec + e * 64 + t) * 16);
                    acc += v.x + v.y;
                }
        }
        // consume: read what landed (the real kernel multiplies it with the key)
        if (mode <= 2 || mode == 5) {
#pragma unroll
            for (int e = 0; e < 8; e++)
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    double2 v = land[(buf * 2 + q) * kSpec + e * 64 + t];
                    acc += v.x + v.y;
                }
        }
    }
    const long long t1 = clock64();
    cluster_sync();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    check[blockIdx.x * 128 + threadIdx.x] = acc;
}

int main() {
    const int iters = 2000;
    long long* d_cyc; double* d_chk;
    const int max_ctas = 296;
    cudaMalloc(&d_cyc, max_ctas * sizeof(long long));
    cudaMalloc(&d_chk, max_ctas * 128 * sizeof(double));
    const size_t smem = (2 + 4) * kSpec * 16 + 64;
    cudaFuncSetAttribute(k_swap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const char* names[] = {"st.async 16 B", "cp.async.bulk 8 KB", "st.shared::cluster + barrier.cluster", "barrier.cluster only", "barrier.cluster + ld.shared::cluster",
                           "half st.async + half cp.async.bulk", "st.async, 8 KB each way (no consume)"};
    printf("{\"iters\": %d, \"bytes_each_way\": %d, \"results\": [\n", iters, 2 * kSpec * 16);
    bool first = true;
    for (int ctas : {2, 148}) {
        for (int mode = 0; mode < 7; mode++) {
            k_swap<<<ctas, 128, smem>>>(mode, iters, d_cyc, d_chk);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { fprintf(stderr, "mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
            std::vector<long long> cyc(ctas);
            std::vector<double> chk(ctas * 128);
            cudaMemcpy(cyc.data(), d_cyc, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
            cudaMemcpy(chk.data(), d_chk, ctas * 128 * sizeof(double), cudaMemcpyDeviceToHost);
            long long mx = 0; for (auto c : cyc) mx = c > mx ? c : mx;
            // expected sum per thread per iteration for modes 0-2, 4: sum over q,e of (peer*10000 + idx) - idx = 16 * peer * 10000
            bool ok = true;
            if (mode != 3 && mode != 6)
                for (int b = 0; b < ctas; b++)
                    for (int th = 0; th < 128; th++)
                        if (chk[b * 128 + th] != 16.0 * ((b & 1) ^ 1) * 10000.0 * iters) ok = false;
            printf("%s {\"ctas\": %d, \"mode\": \"%s\", \"cycles_per_iteration\": %.1f, \"data_ok\": %s}", first ? "" : ",\n", ctas, names[mode],
                   (double)mx / iters, ok ? "true" : "false");
            first = false;
        }
    }
    printf("\n]}\n");
    return 0;
}

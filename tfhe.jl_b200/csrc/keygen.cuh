// keygen.cuh — the steps either side of the gate path on the device: batched LWE
// encryption / phase (api.jl:155-169, lwe.jl:38-59), bootstrapping-key generation (bootstrap.jl:6-15, tgsw.jl:52-88,
// tlwe.jl:63-73) and key-switching-key generation (keyswitch.jl:14-41).
//
// Randomness.  Every kernel that consumes random words takes them from a buffer, so the parity tests feed the oracle
// and the GPU the SAME words and compare bit for bit.  The buffers are filled either by the caller or on the device by
// a counter-based generator (Philox4x32-10, Salmon et al. SC'11: word w of stream (seed, stream) is lane w & 3 of the
// block with counter w >> 2), so a 16 M-gate input set never exists on the host.  Gaussian noise is Box-Muller over
// 53-bit uniforms from the same generator, scaled by sigma and truncated like dtot32 (numeric-functions.jl:51-53).
#pragma once
#include "kernels.cuh"

namespace tfhe_b200 {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x, hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0; k.y += W1;
    }
    return c;
}
__device__ __forceinline__ uint4 philox_block(uint64_t seed, uint64_t stream, uint64_t block) {
    return philox4x32_10(make_uint4((uint32_t)block, (uint32_t)(block >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

// out[i] = word (offset + i) of stream (seed, stream); offset must be a multiple of 4
__global__ void philox_words_kernel(uint32_t* __restrict__ out, size_t count, uint64_t seed, uint64_t stream, uint64_t offset) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 4 >= count) return;
    const uint4 v = philox_block(seed, stream, offset / 4 + b);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; e++)
        if (b * 4 + e < count) out[b * 4 + e] = w[e];
}

// two standard normals per Philox block (Box-Muller on 53-bit uniforms in (0, 1))
__device__ __forceinline__ void philox_normals(uint64_t seed, uint64_t stream, uint64_t block, double& z0, double& z1) {
    const uint4 v = philox_block(seed, stream, block);
    const double u1 = ((double)((((uint64_t)v.x << 32) | v.y) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)((((uint64_t)v.z << 32) | v.w) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = r * c; z1 = r * s;
}
__device__ __forceinline__ int32_t dtot32_dev(double d) { return (int32_t)__double2ll_rz(d * 4294967296.0); }   // numeric-functions.jl:51-53

// out[i] = dtot32(sigma * z_i), z_i = normal i of stream (seed, stream)       (rand_gaussian_torus32, numeric-functions.jl:20-23)
__global__ void gaussian_torus_kernel(int32_t* __restrict__ out, size_t count, double sigma, uint64_t seed, uint64_t stream) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b * 2 >= count) return;
    double z0, z1;
    philox_normals(seed, stream, b, z0, z1);
    out[b * 2] = dtot32_dev(z0 * sigma);
    if (b * 2 + 1 < count) out[b * 2 + 1] = dtot32_dev(z1 * sigma);
}

// keyswitch.jl:28-29: `count` centred noises.  ONE block, fixed summation order: the same seed gives the same key.
__global__ void __launch_bounds__(1024) centred_gaussian_torus_kernel(int32_t* __restrict__ out, double* __restrict__ scratch,
                                                                      int count, double sigma, uint64_t seed, uint64_t stream) {
    __shared__ double part[1024];
    double acc = 0.0;
    for (int b = threadIdx.x; b * 2 < count; b += 1024) {
        double z0, z1;
        philox_normals(seed, stream, (uint64_t)b, z0, z1);
        scratch[b * 2] = z0 * sigma; acc += z0 * sigma;
        if (b * 2 + 1 < count) { scratch[b * 2 + 1] = z1 * sigma; acc += z1 * sigma; }
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    const double mean = part[0] / (double)count;
    for (int i = threadIdx.x; i < count; i += 1024) out[i] = dtot32_dev(scratch[i] - mean);
}

// lwe_encrypt (lwe.jl:38-55), one warp per ciphertext: out[g] = (a[g], mu[g] + noise[g] + <a[g], key>), wrap-around mod 2^32.
// `a` null: the mask words are generated here (stream `a_stream` of `seed`, word g*key_len + i).
// `bits` non-null: mu[g] = bits[g] ? mu_true : -mu_true (api.jl:155-158) instead of mu[g].
__global__ void lwe_encrypt_kernel(const int32_t* __restrict__ key, int key_len, const int32_t* __restrict__ mu,
                                   const unsigned char* __restrict__ bits, int32_t mu_true, const int32_t* __restrict__ noise,
                                   const int32_t* __restrict__ a, uint64_t seed, uint64_t a_stream, int32_t* __restrict__ out,
                                   long long out_stride, size_t count) {
    const size_t g = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= count) return;
    const int lane = threadIdx.x & 31;
    int32_t* o = out + g * out_stride;
    uint32_t acc = 0;
    if (a) {
        const int32_t* ag = a + g * (size_t)key_len;
        for (int i = lane; i < key_len; i += 32) { const int32_t v = ag[i]; o[i] = v; acc += (uint32_t)v * (uint32_t)key[i]; }
    } else {
        const uint64_t w0 = (uint64_t)g * (uint64_t)key_len;
        for (uint64_t blk = w0 / 4 + lane; blk * 4 < w0 + key_len; blk += 32) {
            const uint4 v = philox_block(seed, a_stream, blk);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint64_t idx = blk * 4 + e;
                if (idx >= w0 && idx < w0 + key_len) { const int i = (int)(idx - w0); o[i] = (int32_t)w[e]; acc += w[e] * (uint32_t)key[i]; }
            }
        }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        const uint32_t m = bits ? (bits[g] ? (uint32_t)mu_true : 0u - (uint32_t)mu_true) : (uint32_t)mu[g];
        o[key_len] = (int32_t)(m + (uint32_t)noise[g] + acc);
    }
}

// lwe_phase (lwe.jl:59), one warp per ciphertext: phase[g] = b - <a, key>; bits_out non-null: decrypt (api.jl:167-169)
__global__ void lwe_phase_kernel(const int32_t* __restrict__ key, int key_len, const int32_t* __restrict__ ct, long long ct_stride,
                                 int32_t* __restrict__ phase, unsigned char* __restrict__ bits_out, size_t count) {
    const size_t g = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= count) return;
    const int lane = threadIdx.x & 31;
    const int32_t* c = ct + g * ct_stride;
    uint32_t acc = 0;
    for (int i = lane; i < key_len; i += 32) acc += (uint32_t)c[i] * (uint32_t)key[i];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
        const int32_t ph = (int32_t)((uint32_t)c[key_len] - acc);
        if (phase) phase[g] = ph;
        if (bits_out) bits_out[g] = ph > 0;
    }
}

// tgsw_encrypt (tgsw.jl:84-88) from its parts: sample s = (i, r, j) of the key is the TLWE sample
// (a_s1..a_sk, noise_s + sum_c S_c (*) a_sc) (tlwe.jl:63-73) plus lwe_key[i] * 2^(32 - (r+1)*bgbit) on coefficient 0 of
// component j (tgsw.jl:62-69).  a, prod [S][k][N]; noise [S][N]; bk layout [n][l][k+1][k+1][N].
__global__ void bk_assemble_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ prod, const int32_t* __restrict__ noise,
                                   const int32_t* __restrict__ lwe_key, int32_t* __restrict__ bk, int l, int bgbit, int k) {
    const size_t s = blockIdx.x;                 // sample index (i*l + r)*(k+1) + j
    const int j = (int)(s % (k + 1)), r = (int)((s / (k + 1)) % l), i = (int)((s / (k + 1)) / l);
    const uint32_t g = (uint32_t)lwe_key[i] << (32 - (r + 1) * bgbit);
    int32_t* o = bk + s * (size_t)(k + 1) * kN;
    for (int x = threadIdx.x; x < kN; x += blockDim.x) {
        uint32_t vb = (uint32_t)noise[s * kN + x];
        for (int c = 0; c < k; c++) {
            uint32_t va = (uint32_t)a[(s * k + c) * kN + x];
            vb += (uint32_t)prod[(s * k + c) * kN + x];
            if (x == 0 && j == c) va += g;
            o[c * kN + x] = (int32_t)va;
        }
        if (x == 0 && j == k) vb += g;
        o[k * kN + x] = (int32_t)vb;
    }
}

// keyswitch.jl:35: message(i, j, h) = (in_key[i] * h) << (32 - j*basebit), rows ordered [i][j][h-1]
__global__ void ksk_messages_kernel(const int32_t* __restrict__ in_key, int32_t* __restrict__ msg, int t, int basebit, size_t rows) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int base1 = (1 << basebit) - 1;
    const int h = (int)(row % base1) + 1, j = (int)((row / base1) % t) + 1;
    const size_t i = row / ((size_t)base1 * t);
    msg[row] = (int32_t)(((uint32_t)in_key[i] * (uint32_t)h) << (32 - j * basebit));
}

}  // namespace tfhe_b200

// ---- MK key expansion on the device (RGSW.Expand, mk_internals.jl:304-345; MKBootstrapKey, :447-460) -----------------
namespace tfhe_b200 {

// u[jj][r] = digit r of (b_other[jj] - b_own[jj])   (mk_internals.jl:321, decompose tgsw.jl:99-117); one block per (jj, r)
__global__ void mk_expand_digits_kernel(const int32_t* __restrict__ b_other, const int32_t* __restrict__ b_own,
                                        int32_t* __restrict__ u, int l, int bgbit) {
    const int jj = blockIdx.x / l, r = blockIdx.x % l;
    uint32_t offset = 0;
    for (int q = 1; q <= l; q++) offset += 1u << (32 - q * bgbit);
    offset *= 1u << (bgbit - 1);
    for (int x = threadIdx.x; x < kN; x += blockDim.x) {
        const uint32_t v = (uint32_t)b_other[jj * kN + x] - (uint32_t)b_own[jj * kN + x] + offset;
        u[(size_t)blockIdx.x * kN + x] = (int32_t)((v >> (32 - (r + 1) * bgbit)) & ((1u << bgbit) - 1)) - (1 << (bgbit - 1));
    }
}

// out[j][jj] = base[j][jj] + sum_r u[jj][r] (*) f[j][r]      (mk_internals.jl:327-331 with base = d0, :338 with base = 0)
// U: spectra of the digit polynomials, [l*l][512] (one piece: |u| <= Bg/2); F: spectra of f in two 16-bit pieces,
// [n*l][2][512].  The sum over r is taken in the transform domain: l * (Bg/2) * 2^15 * N <= 2^33 keeps the rounding
// error below 2^-10 (exactness), so lo + (hi << 16) is the exact integer sum of the l products.
// One 64-thread block per (j, jj); the result goes straight into the expanded sample: poly index jj*p + other of the
// x (which = 0) or y (which = 1) block of bk[own][j].
__global__ void __launch_bounds__(64) mk_expand_mac_kernel(const double2* __restrict__ U, const double2* __restrict__ F,
                                                           const int32_t* __restrict__ base, int32_t* __restrict__ bk_own,
                                                           const double2* __restrict__ E, int l, int p, int other, int which) {
    __shared__ double2 X1[512];
    __shared__ double2 X2[kX2Elems];
    const int t = threadIdx.x;
    const int j = blockIdx.x / l, jj = blockIdx.x % l;
    Twiddles w; w.load(E, t);
    double2 lo[8], hi[8];
#pragma unroll
    for (int e = 0; e < 8; e++) { lo[e] = make_double2(0.0, 0.0); hi[e] = make_double2(0.0, 0.0); }
    for (int r = 0; r < l; r++) {
        const double2* u = U + ((size_t)jj * l + r) * kSpectrum + t;
        const double2* f = F + ((size_t)j * l + r) * 2 * kSpectrum + t;
#pragma unroll
        for (int e = 0; e < 8; e++) {
            const double2 uv = __ldg(u + e * 64);
            cmac(lo[e], uv, __ldg(f + e * 64));
            cmac(hi[e], uv, __ldg(f + kSpectrum + e * 64));
        }
    }
    fft512_inverse(lo, w, X1, X2, t, 0);
    fft512_inverse(hi, w, X1, X2, t, 0);
    const size_t spolys = (size_t)l * (2 * p + 2);
    int32_t* o = bk_own + ((size_t)j * spolys + (size_t)which * l * p + (size_t)jj * p + other) * kN;
    const int32_t* b = base ? base + ((size_t)j * l + jj) * kN : nullptr;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int x = t + 64 * m;
        const uint32_t vl = round_to_u32_fast<true>(lo[m].x) + (round_to_u32_fast<true>(hi[m].x) << 16);
        const uint32_t vh = round_to_u32_fast<true>(-lo[m].y) + (round_to_u32_fast<true>(-hi[m].y) << 16);
        o[x] = (int32_t)(vl + (b ? (uint32_t)b[x] : 0u));
        o[x + 512] = (int32_t)(vh + (b ? (uint32_t)b[x + 512] : 0u));
    }
}

// the parts of an expanded sample that are plain copies: x[jj][own] = d0[jj], y[jj][own] = d1[jj] (:327, :336),
// c0, c1 (:341-342).  src [n][l][N]; dst poly index = first + jj*step inside sample j of bk[own].
__global__ void mk_expand_copy_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ bk_own, int l, int p, int first, int step) {
    const int j = blockIdx.x / l, jj = blockIdx.x % l;
    const size_t spolys = (size_t)l * (2 * p + 2);
    const int32_t* s = src + ((size_t)j * l + jj) * kN;
    int32_t* o = bk_own + ((size_t)j * spolys + first + (size_t)jj * step) * kN;
    for (int x = threadIdx.x; x < kN; x += blockDim.x) o[x] = s[x];
}

}  // namespace tfhe_b200

// cabi.cu — host side of libtfhe_b200.so: the C ABI declared in include/tfhe_b200.h.
// No CPU fallback exists in this file: every compute entry point launches the sm_100a kernels of
// kernels.cuh / mk_kernels.cuh or fails with an error code.
#include "../../include/tfhe_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "blind_rotate.cuh"
#include "blind_rotate_lowlat.cuh"
#include "blind_rotate_cluster.cuh"
#include "blind_rotate_wide.cuh"
#include "mk_kernels.cuh"
#include "mk_blind_rotate.cuh"
#include "mk_blind_rotate_lowlat.cuh"
#include "keygen.cuh"

using namespace tfhe_b200;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct tfhe_b200_ctx {
    tfhe_b200_params P{};
    int device = 0;
    uint32_t flags = 0;
    int NP = 2;
    int G = 0;                       // gates per CTA of the blind-rotation kernel (0 = default, see launch_br_np)
    int sm_count = 148;
    int l2_hint = 0;                 // TFHE_B200_L2HINT=1: key chunks fetched with the L2 evict_last policy (K3 with the producer warpgroup)
    int l2_persist = 0;              // TFHE_B200_L2PERSIST=pct: K3 launches carry an access-policy window over the key (hit ratio pct/100, persisting L2 set-aside at its maximum)
    size_t bk_bytes = 0;             // size of d_bk_fft
    int max_clusters = 0;            // two-CTA clusters of the latency kernel the device holds at once (cudaOccupancyMaxActiveClusters)
    int cluster = 1;                 // batches of <= 1 gate per two SMs (two-piece 80-bit set) take the two-CTA cluster kernel; TFHE_B200_CLUSTER=0: off, 2: phase probe
    int lowlat_waves = 2;            // batches of up to this many gates per SM take the latency kernel (TFHE_B200_LOWLAT_WAVES)
    int lowlat = 1;                  // batches of <= 1 gate per SM: one gate per CTA spread over 4 groups + sliced key switch
    int balance_tail = 1;            // K3: the last wave spreads its gates over all SMs (TFHE_B200_BALANCE=0: full CTAs only)
    int ks_tile32 = 1;               // batches that do not fill the SMs with 64-ciphertext tiles take tiles of 32 (TFHE_B200_KS_TILE32=0: off)
    int ks_tile_min = 2560;          // smallest batch that takes the tiled key switch (TFHE_B200_KS_TILE_MIN)
    int ks_tile = 1;                 // large batches: tiled key switch (TFHE_B200_KS_TILE=0: one CTA per ciphertext)
    int mk_ring = 1;                 // MK blind rotation: 1 = TMA key ring, several gates per CTA (mk_blind_rotate.cuh)
    int mk_pw = 1;                   // MK ring kernel: dedicated producer warpgroup (TFHE_B200_MK_PW=0: in-line producer)
    cudaStream_t stream = nullptr;   // used by the host-buffer entry points
    double2* d_E = nullptr;          // exp(-i*pi*x/1024), x < 2048
    double2* d_bk_fft = nullptr;     // single-key: [n][l][2][2][NP][512]; MK: [p][n][l*(2p+2)][NP][512]
    int32_t* d_ksk = nullptr;        // [parties][Nk][t][base-1][stride]
    int ksk_stride = 0;
    bool have_bk = false, have_ksk = false;
    DevBuf bx, by, bz, bout, bu1, bidx, bidx2;
    // second staging slot + copy streams of the host-buffer entry points (host_chunks double-buffers its chunks)
    DevBuf bx2, by2, bz2, bout2;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    // The scratch buffers above are shared by every entry point.  Calls are stream-ordered on the stream they are
    // given, so a call on a DIFFERENT stream than the previous user first waits for that user's event.
    cudaEvent_t scratch_ev = nullptr;
    cudaStream_t scratch_stream = nullptr;
    bool scratch_busy = false;
    std::mutex mu;
    std::atomic<uint64_t> launches{0};
    size_t chunk = 1 << 16;          // gates per host-staged chunk
    size_t dev_piece = (size_t)1 << 20;   // gates per pass of a device-resident batch (bounds the scratch; TFHE_B200_DEV_PIECE)
};

namespace {

// The text of the last error is kept per calling thread (errno style): a failing call on one thread never
// touches a string another thread may be reading through tfhe_b200_last_error.
int fail(tfhe_b200_ctx*, int code, const std::string& msg) {
    g_create_error = msg;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ctx, TFHE_B200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

struct ScratchGuard {
    tfhe_b200_ctx* c; cudaStream_t s;
    ScratchGuard(tfhe_b200_ctx* ctx, cudaStream_t stream) : c(ctx), s(stream) {
        if (c->scratch_busy && c->scratch_stream != s) cudaStreamWaitEvent(s, c->scratch_ev, 0);
    }
    ~ScratchGuard() {
        if (cudaEventRecord(c->scratch_ev, s) == cudaSuccess) { c->scratch_stream = s; c->scratch_busy = true; }
    }
};

int reserve(tfhe_b200_ctx* ctx, DevBuf& b, size_t bytes) {
    if (b.cap >= bytes) return 0;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr; b.cap = 0;
    CU(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return 0;
}

bool single_key_supported(int l, int bgbit) { return (l == 2 && bgbit == 10) || (l == 3 && bgbit == 7); }
bool mk_supported(int p, int l, int bgbit) {
    return (p >= 2 && p <= 8) && ((l == 4 && bgbit == 7) || (l == 5 && bgbit == 6) || (l == 8 && bgbit == 4));
}

int env_int(const char* name, int dflt) {
    const char* s = std::getenv(name);
    return s && *s ? std::atoi(s) : dflt;
}

// ---- kernel dispatch ------------------------------------------------------------------------------
template <int L, int BGBIT, int NP, int G, int STAGES, int MODE, int TM = 0, int OPT = 0>
int launch_br_g(tfhe_b200_ctx* ctx, const BlindRotateArgs& A_in, cudaStream_t s) {
    auto kern = blind_rotate_kernel<L, BGBIT, NP, G, STAGES, MODE, TM, OPT>;
    const size_t smem = br_smem_bytes(NP, G, STAGES, A_in.n_pad, TM);
    if (smem > 227 * 1024) return fail(ctx, TFHE_B200_EINVAL, "LWE dimension too large for the shared-memory layout");
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // gate -> CTA map (BlindRotateArgs::split): full waves of G gates per CTA, then ONE wave that spreads the remaining gates
    // over all SMs (fewer gates per CTA, done sooner) instead of a partial wave of full CTAs
    BlindRotateArgs A = A_in;
    const unsigned long long per_wave = (unsigned long long)G * ctx->sm_count;
    const unsigned long long full_waves = ctx->balance_tail ? (A.count - 1) / per_wave : A.count / per_wave;
    const unsigned long long rest = A.count - full_waves * per_wave;
    A.split = (unsigned)(full_waves * ctx->sm_count);
    A.tail = ctx->balance_tail ? (int)std::max<unsigned long long>(1, (rest + ctx->sm_count - 1) / ctx->sm_count) : G;
    unsigned grid = A.split + (unsigned)((rest + A.tail - 1) / A.tail);
    if (ctx->l2_persist > 0 && ctx->bk_bytes) {
        // experiment: keep the key in the persisting part of L2 instead of re-streaming 40 % of it from DRAM per wave
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeAccessPolicyWindow;
        at[0].val.accessPolicyWindow.base_ptr = (void*)A.bk_fft;
        at[0].val.accessPolicyWindow.num_bytes = ctx->bk_bytes;
        at[0].val.accessPolicyWindow.hitRatio = std::min(1.0f, ctx->l2_persist / 100.0f);
        at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 * G + ((OPT >> 7) & 1) * 128); cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cfg.attrs = at; cfg.numAttrs = 1;
        CU(cudaLaunchKernelEx(&cfg, kern, A));
    } else
    kern<<<grid, 64 * G + ((OPT >> 7) & 1) * 128, smem, s>>>(A);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}
// G = gates resident per CTA (one CTA per SM), STAGES = depth of the TMA key ring.  TFHE_B200_G selects among
// the compiled variants (development knob); the defaults are the fastest measured on B200 (profiles/).
template <int L, int BGBIT, int NP, int MODE>
int launch_br_np(tfhe_b200_ctx* ctx, const BlindRotateArgs& A, cudaStream_t s) {
    if constexpr (MODE == 1) {
        return launch_br_g<L, BGBIT, NP, 2, 2, MODE>(ctx, A, s);
    } else {
        // 4 gates = 8 warps = 2 per SM sub-partition: the only shape that leaves 255 registers per thread
        switch (ctx->G) {
            // development knobs (A/B runs): the round-1 kernel, the output-stationary step without the producer warpgroup
            case 204: return launch_br_g<L, BGBIT, NP, 4, 6, MODE, NP == 2 ? 2 : 0>(ctx, A, s);
            case 304:
                if constexpr (NP == 2) return launch_br_g<L, BGBIT, NP, 4, 5, MODE, 3>(ctx, A, s);
                else return launch_br_g<L, BGBIT, NP, 4, 6, MODE>(ctx, A, s);
            case 384: if constexpr (NP == 2) return launch_br_g<L, BGBIT, NP, 4, 5, MODE, 3, 128>(ctx, A, s); else break;        // default without the paired transforms
            case 1304:                                                                  // clock64 phase probe of the default kernel
                if constexpr (NP == 2 && L == 2 && MODE == 0) {
                    BlindRotateArgs B = A;
                    CU(cudaMalloc(&B.probe, 4 * 8 * 8 * sizeof(unsigned long long)));
                    int rc = launch_br_g<L, BGBIT, NP, 4, 5, MODE, 3, 8 | 128>(ctx, B, s);
                    std::vector<unsigned long long> h(4 * 8 * 8);
                    CU(cudaStreamSynchronize(s));
                    CU(cudaMemcpy(h.data(), B.probe, h.size() * 8, cudaMemcpyDeviceToHost));
                    CU(cudaFree(B.probe));
                    for (int b = 0; b < 2; b++)
                        for (int w = 0; w < 8; w++) {
                            const unsigned long long* o = &h[(b * 8 + w) * 8];
                            fprintf(stderr, "probe cta %d warp %d: total %llu fwd %llu mac %llu inv %llu endbar %llu | key wait %llu first %llu producer %llu (cycles per iteration: %.0f)\n",
                                    b, w, o[0], o[1], o[2], o[3], o[4], o[5], o[6], o[7], (double)o[0] / A.n_iter);
                        }
                    return rc;
                } else return launch_br_g<L, BGBIT, NP, 4, 6, MODE>(ctx, A, s);
            default:
                // Small batches (dependent circuits, single gates): fewer gates per CTA so that every gate gets an
                // SM (sub-partition) of its own — a lone group finishes an iteration ~3x sooner than four sharing
                // the FP64 pipe, and groups without a gate would only burn cycles on zeros.
                {
                    constexpr int TMv = NP == 1 ? 0 : 2;
                    const unsigned long long sms = (unsigned long long)ctx->sm_count;
                    // measured (tools/latency_probe.py, tools/waves_probe.py): one wave of 148 gates takes 1.95 ms on the latency
                    // kernel, two 3.7-3.9 ms, three 5.6-5.8 ms; K3 with a balanced wave 4.0-4.2 ms at 2 gates per CTA, 4.45-4.61 ms
                    // at 3 and 5.67 ms at 4 (all with the key switch): ctx->lowlat_waves
                    if constexpr (NP == 2) {
                        // at most one gate per two SMs: a cluster of two CTAs per gate (blind_rotate_cluster.cuh)
                        if (A.count <= (unsigned long long)ctx->max_clusters && ctx->lowlat && ctx->cluster == 2) {   // clock64 phase probe (development)
                            BlindRotateArgs B = A;
                            CU(cudaMalloc(&B.probe, 80 * sizeof(unsigned long long)));
                            auto kern = blind_rotate_cluster_kernel<L, BGBIT, 1>;
                            const size_t smem = br_cluster_smem_bytes<L>(A.n_pad);
                            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            kern<<<(unsigned)(2 * A.count), br_cluster_threads<L>(), smem, s>>>(B);
                            CU(cudaGetLastError());
                            ctx->launches++;
                            std::vector<unsigned long long> h(80);
                            CU(cudaStreamSynchronize(s));
                            CU(cudaMemcpy(h.data(), B.probe, h.size() * 8, cudaMemcpyDeviceToHost));
                            CU(cudaFree(B.probe));
                            static const char* names[10] = {"rotate", "fwd", "send", "barA", "own", "land", "peer", "barB", "inv", "upd+barC"};
                            for (int w = 0; w < 8; w++) {
                                fprintf(stderr, "cluster probe cta %d warp %d (cycles per iteration):", w / 4, w % 4);
                                for (int k = 0; k < 10; k++) fprintf(stderr, " %s %.0f", names[k], (double)h[w * 10 + k] / A.n_iter);
                                fprintf(stderr, "\n");
                            }
                            return 0;
                        }
                        if (A.count <= (unsigned long long)ctx->max_clusters && ctx->lowlat && ctx->cluster && br_cluster_smem_bytes<L>(A.n_pad) <= 227 * 1024) {
                            auto kern = blind_rotate_cluster_kernel<L, BGBIT>;
                            const size_t smem = br_cluster_smem_bytes<L>(A.n_pad);
                            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                            kern<<<(unsigned)(2 * A.count), br_cluster_threads<L>(), smem, s>>>(A);
                            CU(cudaGetLastError());
                            ctx->launches++;
                            return 0;
                        }
                    }
                    if (A.count <= (unsigned long long)ctx->lowlat_waves * sms && ctx->lowlat && br_lowlat_smem_bytes<L, NP>(A.n_pad) <= 227 * 1024) {
                        // latency path (blind_rotate_lowlat.cuh): one gate per CTA, one digit polynomial per group
                        auto kern = blind_rotate_lowlat_kernel<L, BGBIT, NP>;
                        const size_t smem = br_lowlat_smem_bytes<L, NP>(A.n_pad);
                        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                        kern<<<(unsigned)A.count, 64 * 2 * L, smem, s>>>(A);
                        CU(cudaGetLastError());
                        ctx->launches++;
                        return 0;
                    }
                    if (A.count <= sms) return launch_br_g<L, BGBIT, NP, 1, 6, MODE, TMv>(ctx, A, s);
                    if (A.count <= 2 * sms) return launch_br_g<L, BGBIT, NP, 2, 6, MODE, TMv>(ctx, A, s);
                    // two pieces: output-stationary step + dedicated producer warpgroup + transforms in pairs where l = 2
                    // (profiles/r2/k3_variants.md: 480 vs 563 ms per 65 536 gates)
                    if constexpr (NP == 2) return launch_br_g<L, BGBIT, NP, 4, 5, MODE, 3, 128 | 16>(ctx, A, s);
                    else return launch_br_g<L, BGBIT, NP, 4, 6, MODE, 0, 128>(ctx, A, s);   // one piece: register accumulators + producer warpgroup
                }
        }
        return launch_br_g<L, BGBIT, NP, 4, 6, MODE>(ctx, A, s);   // a two-piece-only variant was asked for with one piece
    }
}
// mask size k > 1: four gates per CTA, output spectra in tensor memory, key from L2
template <int L, int BGBIT, int NP, int MODE>
int launch_br_wide(tfhe_b200_ctx* ctx, const BlindRotateArgs& A_in, cudaStream_t s) {
    const int kp1 = ctx->P.k + 1;
    auto kern = blind_rotate_wide_kernel<L, BGBIT, NP, MODE>;
    const size_t smem = br_wide_smem_bytes(NP, kp1, A_in.n_pad);
    if (smem > 227 * 1024) return fail(ctx, TFHE_B200_EINVAL, "LWE dimension too large for the shared-memory layout");
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BlindRotateArgs A = A_in;   // full waves of kWideGates gates per CTA, then one wave with the rest spread over all SMs
    const unsigned long long per_wave = (unsigned long long)kWideGates * ctx->sm_count;
    const unsigned long long full_waves = (A.count - 1) / per_wave, rest = A.count - full_waves * per_wave;
    A.split = (unsigned)(full_waves * ctx->sm_count);
    A.tail = (int)std::max<unsigned long long>(1, (rest + ctx->sm_count - 1) / ctx->sm_count);
    kern<<<A.split + (unsigned)((rest + A.tail - 1) / A.tail), 64 * kWideGates, smem, s>>>(A, kp1);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}
template <int L, int BGBIT, int NP>
int launch_extern_wide(tfhe_b200_ctx* ctx, const int32_t* acc, const int32_t* idx, int32_t* out, size_t count, cudaStream_t s) {
    const int kp1 = ctx->P.k + 1;
    auto kern = extern_product_wide_kernel<L, BGBIT, NP>;
    const size_t smem = br_wide_group_bytes(NP, kp1, 0, true);
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)count, 64, smem, s>>>(ctx->d_bk_fft, ctx->d_E, acc, idx, out, kp1);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

template <int MODE>
int launch_br(tfhe_b200_ctx* ctx, const BlindRotateArgs& A, cudaStream_t s) {
    const int l = ctx->P.l, bg = ctx->P.bgbit;
    if (A.count == 0) return 0;
    if (ctx->P.k > 1) {   // tlwe_mask_size > 1 (api.jl:30,55): the k-generic kernel of blind_rotate_wide.cuh
        if (l == 2 && bg == 10) return ctx->NP == 2 ? launch_br_wide<2, 10, 2, MODE>(ctx, A, s) : launch_br_wide<2, 10, 1, MODE>(ctx, A, s);
        if (l == 3 && bg == 7) return ctx->NP == 2 ? launch_br_wide<3, 7, 2, MODE>(ctx, A, s) : launch_br_wide<3, 7, 1, MODE>(ctx, A, s);
        return fail(ctx, TFHE_B200_EINVAL, "unsupported (l, bgbit)");
    }
    if (l == 2 && bg == 10) return ctx->NP == 2 ? launch_br_np<2, 10, 2, MODE>(ctx, A, s) : launch_br_np<2, 10, 1, MODE>(ctx, A, s);
    if (l == 3 && bg == 7) return ctx->NP == 2 ? launch_br_np<3, 7, 2, MODE>(ctx, A, s) : launch_br_np<3, 7, 1, MODE>(ctx, A, s);
    return fail(ctx, TFHE_B200_EINVAL, "unsupported (l, bgbit)");
}

int launch_extern(tfhe_b200_ctx* ctx, const int32_t* acc, const int32_t* idx, int32_t* out, size_t count, cudaStream_t s) {
    const int l = ctx->P.l, bg = ctx->P.bgbit;
    if (count == 0) return 0;
    if (ctx->P.k > 1) {
        if (l == 2 && bg == 10) return ctx->NP == 2 ? launch_extern_wide<2, 10, 2>(ctx, acc, idx, out, count, s) : launch_extern_wide<2, 10, 1>(ctx, acc, idx, out, count, s);
        if (l == 3 && bg == 7) return ctx->NP == 2 ? launch_extern_wide<3, 7, 2>(ctx, acc, idx, out, count, s) : launch_extern_wide<3, 7, 1>(ctx, acc, idx, out, count, s);
        return fail(ctx, TFHE_B200_EINVAL, "unsupported (l, bgbit)");
    }
    unsigned grid = (unsigned)count;
    if (l == 2 && bg == 10) {
        if (ctx->NP == 2) extern_product_kernel<2, 10, 2><<<grid, 64, 0, s>>>(ctx->d_bk_fft, ctx->d_E, acc, idx, out);
        else extern_product_kernel<2, 10, 1><<<grid, 64, 0, s>>>(ctx->d_bk_fft, ctx->d_E, acc, idx, out);
    } else if (l == 3 && bg == 7) {
        if (ctx->NP == 2) extern_product_kernel<3, 7, 2><<<grid, 64, 0, s>>>(ctx->d_bk_fft, ctx->d_E, acc, idx, out);
        else extern_product_kernel<3, 7, 1><<<grid, 64, 0, s>>>(ctx->d_bk_fft, ctx->d_E, acc, idx, out);
    } else return fail(ctx, TFHE_B200_EINVAL, "unsupported (l, bgbit)");
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// latency path: several CTAs per ciphertext; the output rows must have been zero-initialised by the caller
int launch_keyswitch_sliced(tfhe_b200_ctx* ctx, const KeyswitchArgs& A, size_t count, cudaStream_t s) {
    // up to 256 CTAs per ciphertext (4 mask positions = 32 independent row loads each): a lone gate's key switch is a chain
    // of L2/HBM round trips, one per position of a slice (1 gate: 70 us with 32 slices, measured under ncu)
    const int slices = (int)std::min<size_t>(256, std::max<size_t>(1, (size_t)16 * ctx->sm_count / count));
    keyswitch_sliced_kernel<<<(unsigned)(count * slices), A.stride / 4, (size_t)(A.Nk / slices + 1) * sizeof(int32_t), s>>>(A, count, slices);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

int launch_keyswitch_args(tfhe_b200_ctx* ctx, const KeyswitchArgs& A, size_t count, cudaStream_t s) {
    // large batches: 64 ciphertexts per CTA, table streamed once per CTA through shared memory (keyswitch_tile_kernel).
    // Measured (tools/ks_probe.py, profiles/r2/ks_probe_v17.json): a wave of 64-ciphertext tiles takes 4.55 ms whatever the
    // batch, a wave of 32-ciphertext tiles 2.75 ms; the per-ciphertext kernel 2.3 ms at 2 048, 3.4 ms at 3 072, 4.2 ms at
    // 4 096 ciphertexts.  So: one CTA per ciphertext below 2 560, tiles of 32 while one wave of them holds the batch
    // (32 x SMs = 4 736), tiles of 64 above.
    if (ctx->ks_tile && A.t == kKsT && A.basebit == kKsBasebit && (A.stride == 512 || A.stride == 640) && count >= (size_t)ctx->ks_tile_min) {
        const size_t smem = ks_tile_smem_bytes(A.stride);
        const bool small = ctx->ks_tile32 && count <= (size_t)32 * ctx->sm_count;
        const unsigned tile = small ? 32 : kKsTile, grid = (unsigned)((count + tile - 1) / tile);
        auto go = [&](auto kern, int threads) -> int {
            CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, threads, smem, s>>>(A, count);
            return 0;
        };
        int rc;
        if (A.stride == 512) rc = small ? go(keyswitch_tile_kernel<512, 1>, 128) : go(keyswitch_tile_kernel<512, 2>, 256);
        else rc = small ? go(keyswitch_tile_kernel<640, 1>, 160) : go(keyswitch_tile_kernel<640, 2>, 320);
        if (rc) return rc;
        CU(cudaGetLastError());
        ctx->launches++;
        return 0;
    }
    keyswitch_kernel<<<(unsigned)count, A.stride / 4, (size_t)A.Nk * sizeof(int32_t), s>>>(A, count);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

// keyswitch of ciphertexts [count][Nk+1] -> [count][n+1] with key set `party`
int launch_keyswitch(tfhe_b200_ctx* ctx, const int32_t* in, int32_t* out, size_t count, cudaStream_t s) {
    if (count == 0) return 0;
    const auto& P = ctx->P;
    KeyswitchArgs A{};
    A.ksk = ctx->d_ksk; A.in = in; A.out = out; A.out_b = out;
    A.n = P.n; A.Nk = P.N * P.k; A.t = P.t; A.basebit = P.basebit; A.stride = ctx->ksk_stride;
    A.in_stride = A.Nk + 1; A.in_offset = 0; A.in_b_offset = A.Nk;
    A.out_stride = P.n + 1; A.out_offset = 0; A.b_offset = P.n; A.b_mode = 0;
    if (count <= 3 * (size_t)ctx->sm_count && ctx->lowlat) {
        CU(cudaMemsetAsync(out, 0, count * (size_t)(P.n + 1) * sizeof(int32_t), s));
        return launch_keyswitch_sliced(ctx, A, count, s);
    }
    return launch_keyswitch_args(ctx, A, count, s);
}

int launch_lincomb(tfhe_b200_ctx* ctx, const int32_t* x, const int32_t* y, int32_t* out, int32_t ka, int32_t kb,
                   int32_t cb, int width, size_t rows, int const_mode, cudaStream_t s) {
    if (rows == 0) return 0;
    unsigned long long total = (unsigned long long)rows * width;
    unsigned grid = (unsigned)((total + 255) / 256);
    lincomb_kernel<<<grid, 256, 0, s>>>(x, y, out, ka, kb, cb, width, total, const_mode);
    CU(cudaGetLastError());
    ctx->launches++;
    return 0;
}

constexpr int32_t kMu8 = 1 << 29;    // encode_message(1, 8)   numeric-functions.jl:42-45
constexpr int32_t kMu4 = 1 << 30;    // encode_message(1, 4)

// prologue coefficients of the ten bootstrapped binary gates (gates.jl)
bool gate_coeffs(int op, int32_t& cb, int32_t& ka, int32_t& kb) {
    switch (op) {
        case TFHE_B200_NAND:  cb = kMu8;  ka = -1; kb = -1; return true;
        case TFHE_B200_OR:    cb = kMu8;  ka = 1;  kb = 1;  return true;
        case TFHE_B200_AND:   cb = -kMu8; ka = 1;  kb = 1;  return true;
        case TFHE_B200_XOR:   cb = kMu4;  ka = 2;  kb = 2;  return true;
        case TFHE_B200_XNOR:  cb = -kMu4; ka = -2; kb = -2; return true;
        case TFHE_B200_NOR:   cb = -kMu8; ka = -1; kb = -1; return true;
        case TFHE_B200_ANDNY: cb = -kMu8; ka = -1; kb = 1;  return true;
        case TFHE_B200_ANDYN: cb = -kMu8; ka = 1;  kb = -1; return true;
        case TFHE_B200_ORNY:  cb = kMu8;  ka = -1; kb = 1;  return true;
        case TFHE_B200_ORYN:  cb = kMu8;  ka = 1;  kb = -1; return true;
        default: return false;
    }
}

BlindRotateArgs br_args(tfhe_b200_ctx* ctx, size_t count) {
    BlindRotateArgs A{};
    A.bk_fft = ctx->d_bk_fft; A.E = ctx->d_E; A.l2_hint = ctx->l2_hint;
    A.n = ctx->P.n; A.n_iter = ctx->P.n; A.n_pad = (ctx->P.n + 3) & ~3;
    A.count = count; A.mu = kMu8;
    return A;
}

int check_single(tfhe_b200_ctx* ctx, bool need_bk, bool need_ksk) {
    if (!ctx) return TFHE_B200_EINVAL;
    if (ctx->P.parties != 1) return fail(ctx, TFHE_B200_EINVAL, "single-key entry point called on an MK context");
    if (need_bk && !ctx->have_bk) return fail(ctx, TFHE_B200_ENOKEY, "bootstrap key not loaded");
    if (need_ksk && !ctx->have_ksk) return fail(ctx, TFHE_B200_ENOKEY, "keyswitch key not loaded");
    CU(cudaSetDevice(ctx->device));
    return 0;
}

// bootstrap_wo_keyswitch with fused prologue, device pointers
int bootstrap_wo_ks_dev(tfhe_b200_ctx* ctx, const int32_t* x, const int32_t* y, int32_t ka, int32_t kb, int32_t cb,
                        int32_t mu, int32_t* out, size_t count, cudaStream_t s) {
    BlindRotateArgs A = br_args(ctx, count);
    A.x = x; A.y = y; A.ka = ka; A.kb = kb; A.cb = cb; A.mu = mu; A.out = out;
    return launch_br<0>(ctx, A, s);
}

// one gate over device-resident ciphertexts; u1 is scratch [count][Nk+1] ([2*count][Nk+1] for MUX)
int gate_dev(tfhe_b200_ctx* ctx, int op, const int32_t* x, const int32_t* y, const int32_t* z, int32_t* out,
             size_t count, int32_t* u1, cudaStream_t s) {
    const int w = ctx->P.n + 1, wu = ctx->P.N * ctx->P.k + 1;
    int32_t cb, ka, kb;
    int rc;
    if (gate_coeffs(op, cb, ka, kb)) {
        if (!x || !y) return fail(ctx, TFHE_B200_EINVAL, "binary gate needs x and y");
        if ((rc = bootstrap_wo_ks_dev(ctx, x, y, ka, kb, cb, kMu8, u1, count, s))) return rc;   // bootstrap.jl:93
        return launch_keyswitch(ctx, u1, out, count, s);                                        // bootstrap.jl:94
    }
    switch (op) {
        case TFHE_B200_NOT:
            if (!x) return fail(ctx, TFHE_B200_EINVAL, "NOT needs x");
            return launch_lincomb(ctx, x, nullptr, out, -1, 0, 0, w, count, 0, s);
        case TFHE_B200_CONSTANT:
            if (!x) return fail(ctx, TFHE_B200_EINVAL, "CONSTANT needs x (value flags)");
            return launch_lincomb(ctx, x, nullptr, out, 0, 0, kMu8, w, count, 1, s);
        case TFHE_B200_MUX:
            if (!x || !y || !z) return fail(ctx, TFHE_B200_EINVAL, "MUX needs x, y and z");
            {   // both bootstraps in ONE launch of 2*count gates: u1 = [AND(x, y) rows | AND(NOT x, z) rows]
                BlindRotateArgs A = br_args(ctx, 2 * count);
                A.x = x; A.y = y; A.ka = 1; A.kb = 1; A.cb = -kMu8;                                   // gates.jl:166-167
                A.x2 = x; A.y2 = z; A.ka2 = -1; A.kb2 = 1; A.cb2 = -kMu8; A.half = count;             // gates.jl:170-171
                A.mu = kMu8; A.out = u1;
                if ((rc = launch_br<0>(ctx, A, s))) return rc;
            }
            if ((rc = launch_lincomb(ctx, u1, u1 + count * (size_t)wu, u1, 1, 1, kMu8, wu, count, 0, s))) return rc;   // gates.jl:174
            return launch_keyswitch(ctx, u1, out, count, s);                                          // gates.jl:176
        default:
            return fail(ctx, TFHE_B200_EINVAL, "unknown gate opcode");
    }
}

}  // namespace

// =====================================================================================================
// All functions below were declared extern "C" by include/tfhe_b200.h and keep that linkage.

int tfhe_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tfhe_b200_create(const tfhe_b200_params* params, int device_id, uint32_t flags, tfhe_b200_ctx** out) {
    tfhe_b200_ctx* ctx = nullptr;   // for the CU macro: errors go to the thread-local string
    if (!params || !out) return fail(nullptr, TFHE_B200_EINVAL, "null argument");
    *out = nullptr;
    const auto& P = *params;
    if (P.N != 1024 || P.k < 1 || P.k > 3) return fail(nullptr, TFHE_B200_EINVAL, "only N = 1024, k = 1..3 are supported");
    if (P.parties != 1 && P.k != 1) return fail(nullptr, TFHE_B200_EINVAL, "MK-TFHE is defined for k = 1 (mk_internals.jl:48)");
    if (P.n < 1 || P.n > 4096) return fail(nullptr, TFHE_B200_EINVAL, "n out of range");
    if (P.t < 1 || P.basebit < 1 || P.t * P.basebit > 31 || P.basebit > 4 || P.n > 4000)
        return fail(nullptr, TFHE_B200_EINVAL, "bad keyswitch parameters (the reference's sets use t = 8, basebit = 2)");
    if (P.parties == 1 ? !single_key_supported(P.l, P.bgbit) : !mk_supported(P.parties, P.l, P.bgbit))
        return fail(nullptr, TFHE_B200_EINVAL, "unsupported (parties, l, bgbit): supported are the reference's parameter sets");
    if (tfhe_b200_device_count() <= device_id || device_id < 0)
        return fail(nullptr, TFHE_B200_ENODEV, "no CUDA device " + std::to_string(device_id) + " (this library has no CPU fallback)");
    CU(cudaSetDevice(device_id));
    tfhe_b200_ctx* c = new tfhe_b200_ctx();
    c->P = P; c->device = device_id; c->flags = flags;
    c->NP = (flags & TFHE_B200_FLAG_UNSPLIT_FFT) ? 1 : 2;
    c->G = env_int("TFHE_B200_G", 0);
    c->mk_ring = env_int("TFHE_B200_MK_RING", 1);
    c->mk_pw = env_int("TFHE_B200_MK_PW", 1);
    c->lowlat = env_int("TFHE_B200_LOWLAT", 1);
    // measured (tools/waves_probe.py, profiles/r2/waves_probe_v17.json): with its last wave balanced K3 takes 4.45-4.61 ms for
    // 297-444 gates (3 per CTA) where three waves of the latency kernel take 5.6-5.8 ms; for 149-296 gates the latency
    // kernel's two waves (3.7-3.9 ms) still beat K3 with 2 gates per CTA (4.0-4.2 ms) with two pieces, not with one (3.0-3.2 vs 2.6-2.8 ms)
    // (80-bit set; the 128-bit set keeps three waves: not measured)
    c->lowlat_waves = std::max(1, env_int("TFHE_B200_LOWLAT_WAVES", P.l == 2 ? (c->NP == 2 ? 2 : 1) : 3));
    c->l2_hint = env_int("TFHE_B200_L2HINT", 0);
    c->l2_persist = env_int("TFHE_B200_L2PERSIST", 0);
    c->cluster = env_int("TFHE_B200_CLUSTER", 1);
    c->ks_tile = env_int("TFHE_B200_KS_TILE", 1);
    c->ks_tile32 = env_int("TFHE_B200_KS_TILE32", 1);
    c->ks_tile_min = std::max(1, env_int("TFHE_B200_KS_TILE_MIN", 2560));
    if (!c->ks_tile32) c->ks_tile_min = std::max(c->ks_tile_min, 4096);
    c->balance_tail = env_int("TFHE_B200_BALANCE", 1);
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device_id) == cudaSuccess) {
            c->sm_count = prop.multiProcessorCount;
            if (c->l2_persist > 0) {
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) != cudaSuccess) { cudaGetLastError(); c->l2_persist = 0; }
                if (env_int("TFHE_B200_VERBOSE", 0))
                    fprintf(stderr, "tfhe_b200: L2 %d MB, persisting max %d MB, window max %d MB\n", prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
            }
        }
    }
    if (c->cluster && c->NP == 2 && P.parties <= 1 && P.k == 1) {
        // how many two-CTA clusters of the latency kernel the device can hold at once (0 on a device or partition that
        // cannot co-schedule CTA pairs with this much shared memory: the one-CTA latency kernel takes over)
        auto probe = [&](auto kern, int threads, size_t smem) {
            int n = 0;
            cudaLaunchConfig_t cfg = {};
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.gridDim = dim3(2 * (unsigned)c->sm_count); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
            return n;
        };
        const int n_pad = (P.n + 31) & ~31;
        if (P.l == 2 && P.bgbit == 10) c->max_clusters = probe(blind_rotate_cluster_kernel<2, 10>, br_cluster_threads<2>(), br_cluster_smem_bytes<2>(n_pad));
        else if (P.l == 3 && P.bgbit == 7) c->max_clusters = probe(blind_rotate_cluster_kernel<3, 7>, br_cluster_threads<3>(), br_cluster_smem_bytes<3>(n_pad));
    }
    {   // gates per host-staged chunk: at least one wave of CTAs, never zero or negative
        const long long v = env_int("TFHE_B200_CHUNK", 1 << 16);
        c->chunk = (size_t)std::max<long long>(v, (long long)4 * c->sm_count);
        c->dev_piece = (size_t)std::max<long long>(env_int("TFHE_B200_DEV_PIECE", 1 << 20), (long long)4 * c->sm_count);
    }
    ctx = c;
    cudaError_t e = cudaEventCreateWithFlags(&c->scratch_ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { tfhe_b200_destroy(c); return fail(nullptr, TFHE_B200_ECUDA, cudaGetErrorString(e)); }
    // twiddle table E[x] = exp(-i*pi*x/1024), computed in long double
    std::vector<double2> E(2048);
    for (int x = 0; x < 2048; x++) {
        long double a = -3.14159265358979323846264338327950288L * (long double)x / 1024.0L;
        E[x] = make_double2((double)cosl(a), (double)sinl(a));
    }
    e = cudaMalloc(&c->d_E, sizeof(double2) * 2048);
    if (e == cudaSuccess) e = cudaMemcpy(c->d_E, E.data(), sizeof(double2) * 2048, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { tfhe_b200_destroy(c); return fail(nullptr, TFHE_B200_ECUDA, cudaGetErrorString(e)); }
    *out = c;
    return 0;
}

void tfhe_b200_destroy(tfhe_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->copy_in) { cudaStreamSynchronize(c->copy_in); cudaStreamDestroy(c->copy_in); }
    if (c->copy_out) { cudaStreamSynchronize(c->copy_out); cudaStreamDestroy(c->copy_out); }
    for (int i = 0; i < 2; i++)
        for (cudaEvent_t ev : {c->ev_in[i], c->ev_done[i], c->ev_out[i]})
            if (ev) cudaEventDestroy(ev);
    if (c->scratch_ev) cudaEventDestroy(c->scratch_ev);
    for (DevBuf* b : {&c->bx, &c->by, &c->bz, &c->bout, &c->bu1, &c->bidx, &c->bidx2, &c->bx2, &c->by2, &c->bz2, &c->bout2})
        if (b->p) cudaFree(b->p);
    if (c->d_E) cudaFree(c->d_E);
    if (c->d_bk_fft) cudaFree(c->d_bk_fft);
    if (c->d_ksk) cudaFree(c->d_ksk);
    delete c;
}

const char* tfhe_b200_last_error(const tfhe_b200_ctx*) { return g_create_error.c_str(); }
uint64_t tfhe_b200_kernel_launches(const tfhe_b200_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int tfhe_b200_synchronize(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    // nothing of this context is in flight any more: the next call need not wait for the previous user of the scratch
    // buffers (this is also what makes a following CUDA-graph capture of *_dev calls legal: no event recorded outside
    // the capture is waited on inside it)
    ctx->scratch_busy = false;
    return 0;
}

// ---- key loading ------------------------------------------------------------------------------------
static int load_bk_polys(tfhe_b200_ctx* ctx, const int32_t* bk, size_t polys) {
    CU(cudaSetDevice(ctx->device));
    if (ctx->d_bk_fft) { CU(cudaFree(ctx->d_bk_fft)); ctx->d_bk_fft = nullptr; }
    CU(cudaMalloc(&ctx->d_bk_fft, polys * ctx->NP * kSpectrum * sizeof(double2)));
    ctx->bk_bytes = polys * ctx->NP * kSpectrum * sizeof(double2);
    // transform in slabs so the int32 staging buffer stays small
    const size_t slab = 1 << 14;
    int32_t* d_tmp = nullptr;
    CU(cudaMalloc(&d_tmp, std::min(slab, polys) * kN * sizeof(int32_t)));
    for (size_t off = 0; off < polys; off += slab) {
        size_t cnt = std::min(slab, polys - off);
        cudaError_t e = cudaMemcpyAsync(d_tmp, bk + off * kN, cnt * kN * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            double2* o = ctx->d_bk_fft + off * ctx->NP * kSpectrum;
            if (ctx->NP == 2) bk_transform_kernel<2><<<(unsigned)cnt, 64, 0, ctx->stream>>>(d_tmp, o, ctx->d_E);
            else bk_transform_kernel<1><<<(unsigned)cnt, 64, 0, ctx->stream>>>(d_tmp, o, ctx->d_E);
            e = cudaGetLastError();
            ctx->launches++;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { cudaFree(d_tmp); return fail(ctx, TFHE_B200_ECUDA, cudaGetErrorString(e)); }
    }
    CU(cudaFree(d_tmp));
    ctx->have_bk = true;
    return 0;
}

int tfhe_b200_load_bk(tfhe_b200_ctx* ctx, const int32_t* bk) {
    if (!ctx || !bk) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->P.parties != 1) return fail(ctx, TFHE_B200_EINVAL, "use tfhe_b200_mk_load_bk on an MK context");
    return load_bk_polys(ctx, bk, (size_t)ctx->P.n * ctx->P.l * (ctx->P.k + 1) * (ctx->P.k + 1));
}

static int load_ksk_sets(tfhe_b200_ctx* ctx, const int32_t* ksk, int sets) {
    const auto& P = ctx->P;
    CU(cudaSetDevice(ctx->device));
    const size_t rows = (size_t)sets * P.N * P.k * P.t * ((1 << P.basebit) - 1);
    const int stride = (P.n + 1 + 31) & ~31;
    if (ctx->d_ksk) { CU(cudaFree(ctx->d_ksk)); ctx->d_ksk = nullptr; }
    CU(cudaMalloc(&ctx->d_ksk, rows * stride * sizeof(int32_t)));
    CU(cudaMemsetAsync(ctx->d_ksk, 0, rows * stride * sizeof(int32_t), ctx->stream));
    CU(cudaMemcpy2DAsync(ctx->d_ksk, stride * sizeof(int32_t), ksk, (P.n + 1) * sizeof(int32_t),
                         (P.n + 1) * sizeof(int32_t), rows, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->ksk_stride = stride;
    ctx->have_ksk = true;
    return 0;
}

int tfhe_b200_load_ksk(tfhe_b200_ctx* ctx, const int32_t* ksk) {
    if (!ctx || !ksk) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->P.parties != 1) return fail(ctx, TFHE_B200_EINVAL, "use tfhe_b200_mk_load_ksk on an MK context");
    return load_ksk_sets(ctx, ksk, 1);
}

// ---- single-key, device buffers ---------------------------------------------------------------------
int tfhe_b200_gate_batch_dev(tfhe_b200_ctx* ctx, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                             int32_t* out, size_t count, void* stream) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, true);
    if (rc) return rc;
    const size_t w = (size_t)ctx->P.n + 1, wu = (size_t)ctx->P.N * ctx->P.k + 1;
    // The extracted samples between blind rotation and key switch (4 KB per gate, 8 KB for MUX) live in a scratch buffer:
    // a device-resident batch of any size is walked in pieces of ~2^20 gates (whole CTA waves) so that the scratch stays
    // below 9 GB however large the batch is (16 M gates would otherwise need 66 GB next to their 96 GB of ciphertexts)
    const size_t wave = (size_t)4 * ctx->sm_count, piece = std::max<size_t>(wave, ctx->dev_piece / wave * wave);
    if ((rc = reserve(ctx, ctx->bu1, (op == TFHE_B200_MUX ? 2 : 1) * std::min(count, piece) * wu * 4))) return rc;
    ScratchGuard guard(ctx, (cudaStream_t)stream);
    if (count == 0) return gate_dev(ctx, op, x, y, z, out, 0, (int32_t*)ctx->bu1.p, (cudaStream_t)stream);   // argument checks only
    for (size_t off = 0; off < count; off += piece) {
        const size_t cnt = std::min(piece, count - off);
        if ((rc = gate_dev(ctx, op, x ? x + off * w : nullptr, y ? y + off * w : nullptr, z ? z + off * w : nullptr, out + off * w, cnt,
                           (int32_t*)ctx->bu1.p, (cudaStream_t)stream))) return rc;
    }
    return 0;
}

int tfhe_b200_bootstrap_wo_ks_batch_dev(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count,
                                        void* stream) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, false);
    if (rc) return rc;
    return bootstrap_wo_ks_dev(ctx, x, nullptr, 1, 0, 0, mu, out, count, (cudaStream_t)stream);
}

int tfhe_b200_keyswitch_batch_dev(tfhe_b200_ctx* ctx, const int32_t* in, int32_t* out, size_t count, void* stream) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, false, true);
    if (rc) return rc;
    return launch_keyswitch(ctx, in, out, count, (cudaStream_t)stream);
}

// ---- single-key, host buffers -----------------------------------------------------------------------
// Generic chunked host driver: up to three inputs of widths win, one output of width wout.
// The chunks are double-buffered: while chunk i computes on ctx->stream, chunk i+1 is copied in on `copy_in` and
// chunk i-1 is copied out on `copy_out` (events order the three streams per staging slot), so for batches of
// more than one chunk the PCIe time hides behind the kernels.  Returns after everything has landed in `hout`.
template <typename F>
static int host_chunks(tfhe_b200_ctx* ctx, const int32_t* const hin[3], const size_t win[3], int32_t* hout, size_t wout,
                       size_t count, F&& body) {
    DevBuf* bin[2][3] = {{&ctx->bx, &ctx->by, &ctx->bz}, {&ctx->bx2, &ctx->by2, &ctx->bz2}};
    DevBuf* bo[2] = {&ctx->bout, &ctx->bout2};
    ScratchGuard guard(ctx, ctx->stream);
    // a batch that fits one chunk is still cut in (up to) 4 pieces of whole CTA waves so that its copies overlap too
    const size_t wave = (size_t)4 * ctx->sm_count;
    size_t chunk = ctx->chunk;
    // ... but never in more pieces than the tiled key switch has waves of CTAs (a wave of 64-ciphertext tiles takes 4.55 ms however
    // few tiles it holds): 16 384 gates go in 2 pieces, not 4 (4 launches of 65 tiles cost 18 ms, 2 of 130 cost 9)
    if (count <= chunk && count >= 8 * wave) {
        const size_t ks_wave = (size_t)kKsTile * ctx->sm_count;
        const size_t pieces = std::min<size_t>(4, std::max<size_t>(1, (count + ks_wave - 1) / ks_wave));
        chunk = ((count + pieces - 1) / pieces + wave - 1) / wave * wave;
    }
    bool used[2] = {false, false};
    // The copy-out of chunk i is issued only after the kernels of chunk i+1 have been queued.  The caller's buffers
    // may be PAGEABLE (a Julia Matrix{Int32}, a plain numpy array): cudaMemcpyAsync then blocks the host until the
    // copy has been staged (H2D) or has finished (D2H).  Issued in this order, the host blocks while the device is
    // busy with the next chunk, so the pipeline keeps its overlap with pageable memory too.
    size_t prev_off = 0, prev_cnt = 0; int prev_slot = -1;
    auto copy_out = [&](int slot, size_t off, size_t cnt) -> int {
        CU(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_done[slot], 0));
        CU(cudaMemcpyAsync(hout + off * wout, bo[slot]->p, cnt * wout * 4, cudaMemcpyDeviceToHost, ctx->copy_out));
        CU(cudaEventRecord(ctx->ev_out[slot], ctx->copy_out));
        return 0;
    };
    size_t idx = 0;
    for (size_t off = 0; off < count; off += chunk, idx++) {
        const int slot = (int)(idx & 1);
        const size_t cnt = std::min(chunk, count - off);
        int rc;
        int32_t* din[3] = {nullptr, nullptr, nullptr};
        if (used[slot]) CU(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[slot], 0));   // inputs of chunk idx-2 consumed
        for (int a = 0; a < 3; a++) {
            if (!hin[a]) continue;
            if ((rc = reserve(ctx, *bin[slot][a], std::min(chunk, count) * win[a] * 4))) return rc;
            din[a] = (int32_t*)bin[slot][a]->p;
            CU(cudaMemcpyAsync(din[a], hin[a] + off * win[a], cnt * win[a] * 4, cudaMemcpyHostToDevice, ctx->copy_in));
        }
        CU(cudaEventRecord(ctx->ev_in[slot], ctx->copy_in));
        if ((rc = reserve(ctx, *bo[slot], std::min(chunk, count) * wout * 4))) return rc;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0));
        if (used[slot]) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[slot], 0));     // output of chunk idx-2 copied out
        if ((rc = body(din[0], din[1], din[2], (int32_t*)bo[slot]->p, cnt))) return rc;
        CU(cudaEventRecord(ctx->ev_done[slot], ctx->stream));
        if (prev_slot >= 0 && (rc = copy_out(prev_slot, prev_off, prev_cnt))) return rc;
        prev_slot = slot; prev_off = off; prev_cnt = cnt;
        used[slot] = true;
    }
    if (prev_slot >= 0) { int rc = copy_out(prev_slot, prev_off, prev_cnt); if (rc) return rc; }
    CU(cudaStreamSynchronize(ctx->copy_out));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int tfhe_b200_gate_batch(tfhe_b200_ctx* ctx, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                         int32_t* out, size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, true);
    if (rc) return rc;
    if (!out) return fail(ctx, TFHE_B200_EINVAL, "null output");
    const size_t w = (size_t)ctx->P.n + 1, wu = (size_t)ctx->P.N * ctx->P.k + 1;
    const int32_t* hin[3] = {x, y, z};
    const size_t win[3] = {w, w, w};
    return host_chunks(ctx, hin, win, out, w, count, [&](int32_t* dx, int32_t* dy, int32_t* dz, int32_t* dout, size_t cnt) {
        int r;
        if ((r = reserve(ctx, ctx->bu1, (op == TFHE_B200_MUX ? 2 : 1) * cnt * wu * 4))) return r;
        return gate_dev(ctx, op, dx, dy, dz, dout, cnt, (int32_t*)ctx->bu1.p, ctx->stream);
    });
}

int tfhe_b200_bootstrap_wo_ks_batch(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, false);
    if (rc) return rc;
    if (!x || !out) return fail(ctx, TFHE_B200_EINVAL, "null argument");
    const size_t w = (size_t)ctx->P.n + 1, wu = (size_t)ctx->P.N * ctx->P.k + 1;
    const int32_t* hin[3] = {x, nullptr, nullptr};
    const size_t win[3] = {w, 0, 0};
    return host_chunks(ctx, hin, win, out, wu, count, [&](int32_t* dx, int32_t*, int32_t*, int32_t* dout, size_t cnt) {
        return bootstrap_wo_ks_dev(ctx, dx, nullptr, 1, 0, 0, mu, dout, cnt, ctx->stream);
    });
}

int tfhe_b200_keyswitch_batch(tfhe_b200_ctx* ctx, const int32_t* in, int32_t* out, size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, false, true);
    if (rc) return rc;
    if (!in || !out) return fail(ctx, TFHE_B200_EINVAL, "null argument");
    const size_t w = (size_t)ctx->P.n + 1, wu = (size_t)ctx->P.N * ctx->P.k + 1;
    const int32_t* hin[3] = {in, nullptr, nullptr};
    const size_t win[3] = {wu, 0, 0};
    return host_chunks(ctx, hin, win, out, w, count, [&](int32_t* din, int32_t*, int32_t*, int32_t* dout, size_t cnt) {
        return launch_keyswitch(ctx, din, dout, cnt, ctx->stream);
    });
}

int tfhe_b200_bootstrap_batch(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, true);
    if (rc) return rc;
    if (!x || !out) return fail(ctx, TFHE_B200_EINVAL, "null argument");
    const size_t w = (size_t)ctx->P.n + 1, wu = (size_t)ctx->P.N * ctx->P.k + 1;
    const int32_t* hin[3] = {x, nullptr, nullptr};
    const size_t win[3] = {w, 0, 0};
    return host_chunks(ctx, hin, win, out, w, count, [&](int32_t* dx, int32_t*, int32_t*, int32_t* dout, size_t cnt) {
        int r;
        if ((r = reserve(ctx, ctx->bu1, cnt * wu * 4))) return r;
        if ((r = bootstrap_wo_ks_dev(ctx, dx, nullptr, 1, 0, 0, mu, (int32_t*)ctx->bu1.p, cnt, ctx->stream))) return r;
        return launch_keyswitch(ctx, (int32_t*)ctx->bu1.p, dout, cnt, ctx->stream);
    });
}

int tfhe_b200_extern_product_batch(tfhe_b200_ctx* ctx, const int32_t* acc, const int32_t* bk_index, int32_t* out,
                                   size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, false);
    if (rc) return rc;
    if (!acc || !bk_index || !out) return fail(ctx, TFHE_B200_EINVAL, "null argument");
    for (size_t g = 0; g < count; g++)
        if (bk_index[g] < 0 || bk_index[g] >= ctx->P.n) return fail(ctx, TFHE_B200_EINVAL, "bk_index out of range");
    const int32_t* hin[3] = {acc, bk_index, nullptr};
    const size_t wacc = (size_t)(ctx->P.k + 1) * kN;
    const size_t win[3] = {wacc, 1, 0};
    return host_chunks(ctx, hin, win, out, wacc, count, [&](int32_t* dacc, int32_t* didx, int32_t*, int32_t* dout, size_t cnt) {
        return launch_extern(ctx, dacc, didx, dout, cnt, ctx->stream);
    });
}

int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx* ctx, const int32_t* acc_in, const int32_t* bara, int32_t n_iter,
                                 int32_t* acc_out, size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_single(ctx, true, false);
    if (rc) return rc;
    if (!acc_in || !bara || !acc_out) return fail(ctx, TFHE_B200_EINVAL, "null argument");
    if (n_iter < 0 || n_iter > ctx->P.n) return fail(ctx, TFHE_B200_EINVAL, "n_iter out of range");
    const int32_t* hin[3] = {acc_in, bara, nullptr};
    const size_t wacc = (size_t)(ctx->P.k + 1) * kN;
    const size_t win[3] = {wacc, (size_t)ctx->P.n, 0};
    return host_chunks(ctx, hin, win, acc_out, wacc, count, [&](int32_t* dacc, int32_t* dbara, int32_t*, int32_t* dout, size_t cnt) {
        BlindRotateArgs A = br_args(ctx, cnt);
        A.acc_in = dacc; A.bara_in = dbara; A.n_iter = n_iter; A.out = dout;
        return launch_br<1>(ctx, A, ctx->stream);
    });
}

int tfhe_b200_polymul_batch(tfhe_b200_ctx* ctx, const int32_t* x, const int32_t* y, int32_t* out, size_t count) {
    if (!ctx) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!x || !y || !out) return fail(ctx, TFHE_B200_EINVAL, "null argument");
    CU(cudaSetDevice(ctx->device));
    const int32_t* hin[3] = {x, y, nullptr};
    const size_t win[3] = {(size_t)kN, (size_t)kN, 0};
    return host_chunks(ctx, hin, win, out, (size_t)kN, count, [&](int32_t* dx, int32_t* dy, int32_t*, int32_t* dout, size_t cnt) {
        polymul_kernel<<<(unsigned)cnt, 64, 0, ctx->stream>>>(dx, dy, dout, ctx->d_E);
        CU(cudaGetLastError());
        ctx->launches++;
        return 0;
    });
}

#include "mk_cabi.inc"
#include "keygen_cabi.inc"

// ---- measurement helper -------------------------------------------------------------------------------
int tfhe_b200_measure_fp64_tflops(tfhe_b200_ctx* ctx, double* out_tflops) {
    if (!ctx || !out_tflops) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 8192;
    int rc;
    if ((rc = reserve(ctx, ctx->bout, (size_t)blocks * threads * sizeof(double)))) return rc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->bout.p, 1.0000001, 1e-9, iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double tf = (double)blocks * threads * iters * 8 * 2 / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *out_tflops = best;
    return 0;
}

int tfhe_b200_measure_lds_gbps(tfhe_b200_ctx* ctx, double* out_gbps) {
    if (!ctx || !out_gbps) return TFHE_B200_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512, iters = 4096;
    int rc;
    if ((rc = reserve(ctx, ctx->bout, (size_t)blocks * threads * sizeof(double)))) return rc;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0, ctx->stream));
        lds_peak_kernel<<<blocks, threads, 0, ctx->stream>>>((double*)ctx->bout.p, iters);
        CU(cudaEventRecord(e1, ctx->stream));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        double gbps = (double)blocks * threads * iters * 8 * 16 / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbps > best) best = gbps;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *out_gbps = best;
    return 0;
}

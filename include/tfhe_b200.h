/*
 * tfhe_b200.h — C ABI of libtfhe_b200.so, the B200-native gate-bootstrapping engine.
 *
 * This is the drop-in boundary for ONE hot path of nucypher/TFHE.jl: bootstrapped gate
 * evaluation (linear prologue -> modulus switch -> blind rotation -> sample extraction ->
 * key switch), single-key and multi-key.  The reference (pure Julia) has no FFI; the seam is
 * the set of Julia functions that gates.jl / mk_gates.jl call down into.  Each entry point
 * below names the reference function it replaces (file:line under /root/reference/src).
 * The Julia-side `ccall` binding is shown in INTEGRATION.md and shipped in
 * tfhe.jl_b200/julia/TFHEB200.jl; tests bind the same symbols with Python ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative TFHE_B200_E* code otherwise; the text of
 *     the last error is available from tfhe_b200_last_error() (no exceptions cross the ABI).
 *   - all torus data are int32 (Torus32 = Int32, numeric-functions.jl:1), wrap-around mod 2^32.
 *   - the caller owns every host buffer; the library owns device memory.  Host pointers may be
 *     pageable.  *_dev variants take DEVICE pointers (inputs already resident in HBM) and a
 *     cudaStream_t passed as void* (NULL = default stream); they are asynchronous and stream-ordered:
 *     the library's scratch buffers are shared by all calls of a context, and a call given a different
 *     stream than the previous one first waits (on the device) for that call's work.
 *   - the kernel shape is chosen per call from `count`: up to one gate per two SMs a cluster of two CTAs per gate,
 *     up to 2 gates per SM a latency kernel (one gate per CTA) and a sliced key switch, above that up to 4 gates per CTA
 *     (the last wave of CTAs carries fewer), from 2 560 ciphertexts a tiled key switch; the results do not depend on
 *     the choice.
 *   - there is NO CPU fallback: without a CUDA device every call fails with TFHE_B200_ENODEV.
 *   - calls on one context are serialised by an internal mutex; contexts are independent
 *     (one context per GPU; gate batches shard across contexts with no collective: tfhe_b200_multi_* below).
 *   - error texts are kept per calling thread (errno style).
 *
 * Layouts
 *   LWE ciphertext batch     [count][n+1]            a[0..n-1], b   (lwe.jl:21-29); this is the
 *                                                    memory of a Julia Matrix{Int32}(n+1, count)
 *   extracted LWE batch      [count][N*k+1]          (tlwe.jl:55-59)
 *   BK, coefficient domain   [n][l][k+1][k+1][N]     samples[r,j].a[c] (tgsw.jl:28), int32
 *   KSK                      [N*k][t][base-1][n+1]   key[h,j,i] (keyswitch.jl:36-38), C order of
 *                                                    Julia's column-major array, samples flattened
 *   TLWE batch               [count][k+1][N]         (tlwe.jl:34-41)
 *   MK LWE ciphertext batch  [count][p*n+1]          a[party][n] (mk_internals.jl:9), then b
 *   MK extracted LWE batch   [count][p*N+1]
 *   MK TLWE batch            [count][p+1][N]         a_1..a_p, b (mk_internals.jl:46-57)
 *   MK BK, coefficient dom.  [p][n] samples (bk.key[j,i], mk_internals.jl:453-455), each
 *                            x[l][p][N] | y[l][p][N] | c0[l][N] | c1[l][N]  (:240-250)
 *   MK KSK                   [p] single-key KSKs    (mk_api.jl:66-71)
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFHE_B200_OK 0
#define TFHE_B200_EINVAL (-1)    /* bad argument / unsupported parameter set */
#define TFHE_B200_ENODEV (-2)    /* no usable CUDA device */
#define TFHE_B200_ECUDA (-3)     /* CUDA runtime error (see last_error) */
#define TFHE_B200_ENOKEY (-4)    /* bootstrap / keyswitch key not loaded */
#define TFHE_B200_ENOMEM (-5)

/* SchemeParameters (api.jl:4-21), integer part; noise parameters never cross the ABI. */
typedef struct {
    int32_t n;        /* lwe_size */
    int32_t N;        /* tlwe_polynomial_degree (must be 1024) */
    int32_t k;        /* tlwe_mask_size (1; 2 and 3 on single-key contexts: blind_rotate_wide.cuh) */
    int32_t l;        /* bs_decomp_length */
    int32_t bgbit;    /* bs_log2_base */
    int32_t t;        /* ks_decomp_length */
    int32_t basebit;  /* ks_log2_base */
    int32_t parties;  /* 1 = single-key context; 2..8 = MK-TFHE context (mk_api.jl:4-34) */
} tfhe_b200_params;

/* gate opcodes, in the order of gates.jl */
enum {
    TFHE_B200_NAND = 0,      /* gate_nand     gates.jl:15-18   */
    TFHE_B200_OR = 1,        /* gate_or       gates.jl:27-30   */
    TFHE_B200_AND = 2,       /* gate_and      gates.jl:39-42   */
    TFHE_B200_XOR = 3,       /* gate_xor      gates.jl:51-54   */
    TFHE_B200_XNOR = 4,      /* gate_xnor     gates.jl:63-66   */
    TFHE_B200_NOT = 5,       /* gate_not      gates.jl:76-79   */
    TFHE_B200_CONSTANT = 6,  /* gate_constant gates.jl:91-93   (x[g*(n+1)] != 0 selects true) */
    TFHE_B200_NOR = 7,       /* gate_nor      gates.jl:102-105 */
    TFHE_B200_ANDNY = 8,     /* gate_andny    gates.jl:114-117 */
    TFHE_B200_ANDYN = 9,     /* gate_andyn    gates.jl:126-129 */
    TFHE_B200_ORNY = 10,     /* gate_orny     gates.jl:138-141 */
    TFHE_B200_ORYN = 11,     /* gate_oryn     gates.jl:150-153 */
    TFHE_B200_MUX = 12       /* gate_mux      gates.jl:163-177 */
};

/* create flags */
#define TFHE_B200_FLAG_SPLIT_FFT 0u    /* default: torus operand split in 16-bit halves; rounding bound proven */
#define TFHE_B200_FLAG_UNSPLIT_FFT 1u  /* the reference's own precision regime (polynomials.jl:138-140): one 32-bit piece */

typedef struct tfhe_b200_ctx tfhe_b200_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int tfhe_b200_device_count(void);
/* One context = one parameter set + one key set on one GPU.  Replaces the role of CloudKey
 * (api.jl:111-127) / MKCloudKey (mk_api.jl:85-101) as the holder of evaluation keys. */
int tfhe_b200_create(const tfhe_b200_params* params, int device_id, uint32_t flags, tfhe_b200_ctx** out);
void tfhe_b200_destroy(tfhe_b200_ctx* ctx);
/* ctx may be NULL: returns the last error of a failed create on this thread. */
const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
uint64_t tfhe_b200_kernel_launches(const tfhe_b200_ctx* ctx);
int tfhe_b200_synchronize(tfhe_b200_ctx* ctx);
/* measured FP64 FMA throughput of the context's device (roofline denominator of the transform kernels) */
int tfhe_b200_measure_fp64_tflops(tfhe_b200_ctx* ctx, double* out_tflops);
/* measured conflict-free LDS.128 read rate of the context's device (roofline denominator of the tiled key switch) */
int tfhe_b200_measure_lds_gbps(tfhe_b200_ctx* ctx, double* out_gbps);

/* ---- key loading (K6) -------------------------------------------------------------------- */
/* BootstrapKey (bootstrap.jl:1-16): takes the int32 coefficient form and performs
 * forward_transform.(bk) (bootstrap.jl:12 -> tgsw.jl:120-121) on the device. */
int tfhe_b200_load_bk(tfhe_b200_ctx* ctx, const int32_t* bk);
/* KeyswitchKey.key (keyswitch.jl:7-13,35-38) */
int tfhe_b200_load_ksk(tfhe_b200_ctx* ctx, const int32_t* ksk);

/* ---- single-key hot path, host buffers ---------------------------------------------------- */
/* gate_* (gates.jl:15-177) over a batch; unused inputs may be NULL. */
int tfhe_b200_gate_batch(tfhe_b200_ctx* ctx, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                         int32_t* out, size_t count);
/* bootstrap (bootstrap.jl:92-95) */
int tfhe_b200_bootstrap_batch(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count);
/* bootstrap_wo_keyswitch (bootstrap.jl:69-82): out is [count][N*k+1] */
int tfhe_b200_bootstrap_wo_ks_batch(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count);
/* keyswitch (keyswitch.jl:45-80): in is [count][N*k+1] */
int tfhe_b200_keyswitch_batch(tfhe_b200_ctx* ctx, const int32_t* in, int32_t* out, size_t count);
/* tgsw_extern_mul (tgsw.jl:125-129): out[g] = BK[bk_index[g]] (x) acc[g]; acc/out [count][k+1][N] */
int tfhe_b200_extern_product_batch(tfhe_b200_ctx* ctx, const int32_t* acc, const int32_t* bk_index,
                                   int32_t* out, size_t count);
/* blind_rotate (bootstrap.jl:32-39) on explicit accumulators: acc in/out [count][k+1][N],
 * bara [count][n] already modulus-switched; only the first n_iter key elements are applied. */
int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx* ctx, const int32_t* acc_in, const int32_t* bara, int32_t n_iter,
                                 int32_t* acc_out, size_t count);
/* transformed_mul (polynomials.jl:142-144), exact for ANY int32 operands: x,y,out [count][N] */
int tfhe_b200_polymul_batch(tfhe_b200_ctx* ctx, const int32_t* x, const int32_t* y, int32_t* out, size_t count);

/* ---- single-key hot path, device buffers (asynchronous on `stream`) ------------------------ */
/* A batch of any size: it is walked in pieces of ~2^20 gates (whole CTA waves) so that the library's scratch (the
 * extracted samples between blind rotation and key switch, 4 KB per gate) stays below 9 GB of the 180 GB. */
int tfhe_b200_gate_batch_dev(tfhe_b200_ctx* ctx, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                             int32_t* out, size_t count, void* stream);
int tfhe_b200_bootstrap_wo_ks_batch_dev(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out,
                                        size_t count, void* stream);
int tfhe_b200_keyswitch_batch_dev(tfhe_b200_ctx* ctx, const int32_t* in, int32_t* out, size_t count, void* stream);

/* ---- multi-key hot path (context created with params.parties >= 2) ------------------------- */
/* MKBootstrapKey.key (mk_internals.jl:442-461), int32 coefficient form */
int tfhe_b200_mk_load_bk(tfhe_b200_ctx* ctx, const int32_t* mk_bk);
/* MKCloudKey.keyswitch_key (mk_api.jl:89,97): parties KSKs back to back */
int tfhe_b200_mk_load_ksk(tfhe_b200_ctx* ctx, const int32_t* mk_ksk);
/* mk_gate_nand (mk_gates.jl:7-12) */
int tfhe_b200_mk_nand_batch(tfhe_b200_ctx* ctx, const int32_t* x, const int32_t* y, int32_t* out, size_t count);
int tfhe_b200_mk_nand_batch_dev(tfhe_b200_ctx* ctx, const int32_t* x, const int32_t* y, int32_t* out, size_t count,
                                void* stream);
/* mk_bootstrap (mk_internals.jl:512-515): mk_keyswitch(ks, mk_bootstrap_wo_keyswitch(bk, mu, x)), [count][p*n+1] in and out */
int tfhe_b200_mk_bootstrap_batch(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count);
int tfhe_b200_mk_bootstrap_batch_dev(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count,
                                     void* stream);
/* mk_bootstrap_wo_keyswitch (mk_internals.jl:498-509): out [count][p*N+1] */
int tfhe_b200_mk_bootstrap_wo_ks_batch(tfhe_b200_ctx* ctx, int32_t mu, const int32_t* x, int32_t* out, size_t count);
/* mk_keyswitch (mk_internals.jl:397-411): in [count][p*N+1] -> out [count][p*n+1] */
int tfhe_b200_mk_keyswitch_batch(tfhe_b200_ctx* ctx, const int32_t* in, int32_t* out, size_t count);
/* mk_tgsw_extern_mul (mk_internals.jl:348-391): out[g] = BK[party[g]][bk_index[g]] (x) acc[g]; acc/out [count][p+1][N] */
int tfhe_b200_mk_extern_product_batch(tfhe_b200_ctx* ctx, const int32_t* acc, const int32_t* party,
                                      const int32_t* bk_index, int32_t* out, size_t count);

/* ---- the steps either side of the path on the device ---------------------- */
/* Every function here that consumes randomness exists in two forms: *_words takes the random words from the caller
 * (so a CPU oracle fed the same words must produce the same bits), the other generates them on the device with a
 * counter-based generator (Philox4x32-10; mask words = stream 1 of `seed`, Gaussian noise = Box-Muller over stream 2)
 * and never moves them through the host.  key_len is n for single-key and p*n for MK ciphertexts. */
/* word i of Philox stream (seed, stream), i < count (tests, reproducibility) */
int tfhe_b200_random_words(tfhe_b200_ctx* ctx, uint64_t seed, uint64_t stream, int32_t* out, size_t count);
/* lwe_encrypt (lwe.jl:38-55): out[g] = (a[g], mu[g] + noise[g] + <a[g], key>); a [count][key_len], out [count][key_len+1] */
int tfhe_b200_lwe_encrypt_words_batch(tfhe_b200_ctx* ctx, const int32_t* key, int32_t key_len, const int32_t* mu,
                                      const int32_t* noise, const int32_t* a, int32_t* out, size_t count);
/* encrypt (api.jl:155-158) / mk_encrypt (mk_api.jl:110-126) of count bits (bytes, 0 or 1) with noise stddev sigma */
int tfhe_b200_encrypt_batch(tfhe_b200_ctx* ctx, const int32_t* key, int32_t key_len, const uint8_t* bits, double sigma,
                            uint64_t seed, int32_t* out, size_t count);
/* the same with the ciphertexts left in HBM (out_dev is a DEVICE pointer; key and bits are host pointers) */
int tfhe_b200_encrypt_batch_dev(tfhe_b200_ctx* ctx, const int32_t* key, int32_t key_len, const uint8_t* bits, double sigma,
                                uint64_t seed, int32_t* out_dev, size_t count, void* stream);
/* lwe_phase (lwe.jl:59) and decrypt (api.jl:167-169, mk_api.jl:135-138): phase[g] = b - <a, key>, bits_out[g] = phase > 0;
 * either output may be NULL */
int tfhe_b200_lwe_phase_batch(tfhe_b200_ctx* ctx, const int32_t* key, int32_t key_len, const int32_t* ct, int32_t* phase,
                              uint8_t* bits_out, size_t count);
/* BootstrapKey (bootstrap.jl:6-15; tgsw_encrypt tgsw.jl:84-88, tlwe_encrypt_zero tlwe.jl:63-73) generated, transformed
 * and loaded on the device: lwe_key [n], tlwe_key [k][N], a [n*l*(k+1)][k][N], noise [n*l*(k+1)][N] (sample (i, r, j):
 * its k mask polynomials and its noise polynomial).  bk_out (nullable) receives the coefficient form [n][l][k+1][k+1][N]. */
int tfhe_b200_keygen_bk_words(tfhe_b200_ctx* ctx, const int32_t* lwe_key, const int32_t* tlwe_key, const int32_t* a,
                              const int32_t* noise, int32_t* bk_out);
int tfhe_b200_keygen_bk(tfhe_b200_ctx* ctx, const int32_t* lwe_key, const int32_t* tlwe_key, double sigma, uint64_t seed,
                        int32_t* bk_out);
/* KeyswitchKey (keyswitch.jl:14-41) generated and loaded on the device: out_key [n], in_key [N*k], a [N*k][t][base-1][n],
 * noise [N*k][t][base-1] (already centred, keyswitch.jl:29).  ksk_out (nullable) receives [N*k][t][base-1][n+1]. */
int tfhe_b200_keygen_ksk_words(tfhe_b200_ctx* ctx, const int32_t* out_key, const int32_t* in_key, const int32_t* a,
                               const int32_t* noise, int32_t* ksk_out);
int tfhe_b200_keygen_ksk(tfhe_b200_ctx* ctx, const int32_t* out_key, const int32_t* in_key, double sigma, uint64_t seed,
                         int32_t* ksk_out);

/* MKBootstrapKey (mk_internals.jl:442-461): RGSW.Expand (:304-345) of every party's uni-encryptions — the
 * n*p*(p-1)*2*l^2 polynomial products, their sums, the transform and the load — on the device, no host round trip.
 * uni_enc [p][6][n][l][N] in the order c0, c1, d0, d1, f0, f1 (mk_tgsw_encrypt, :185-227), public_b [p][l][N]
 * (PublicKey.b, :115-139); bk_out (nullable) receives the coefficient form [p][n][l*(2p+2)][N]. */
int tfhe_b200_mk_expand_load_bk(tfhe_b200_ctx* ctx, const int32_t* uni_enc, const int32_t* public_b, int32_t* bk_out);

/* ---- one logical context over several GPUs ------------------------------------ */
/* Gates are independent (gates.jl:15-18) and the evaluation keys are read-only, so a batch shards across GPUs with
 * no exchange step: keys are replicated at load, every call cuts its batch into contiguous shards of whole CTA
 * waves, one host thread per device runs its shard through the single-device entry point above and writes a
 * disjoint slice of the caller's output.  device_ids == NULL or n_dev <= 0 selects every visible device.
 * This is what a Julia `CloudKey(...; devices = ...)` holds so that ONE gate_nand.(ck, xs, ys) uses every GPU. */
typedef struct tfhe_b200_multi tfhe_b200_multi;
int tfhe_b200_multi_create(const tfhe_b200_params* params, const int* device_ids, int n_dev, uint32_t flags,
                           tfhe_b200_multi** out);
void tfhe_b200_multi_destroy(tfhe_b200_multi* m);
/* last error of a tfhe_b200_multi_* call on the calling thread (m may be NULL) */
const char* tfhe_b200_multi_last_error(const tfhe_b200_multi* m);
int tfhe_b200_multi_devices(const tfhe_b200_multi* m);
/* the single-device context of the i-th listed device (for the *_dev entry points); owned by m */
tfhe_b200_ctx* tfhe_b200_multi_context(tfhe_b200_multi* m, int i);
uint64_t tfhe_b200_multi_kernel_launches(const tfhe_b200_multi* m);
/* key replication: same layouts as the single-device loaders, one host thread per device */
int tfhe_b200_multi_load_bk(tfhe_b200_multi* m, const int32_t* bk);
int tfhe_b200_multi_load_ksk(tfhe_b200_multi* m, const int32_t* ksk);
int tfhe_b200_multi_mk_load_bk(tfhe_b200_multi* m, const int32_t* mk_bk);
int tfhe_b200_multi_mk_load_ksk(tfhe_b200_multi* m, const int32_t* mk_ksk);
/* gate_* (gates.jl:15-177), bootstrap (bootstrap.jl:92-95) / mk_bootstrap (mk_internals.jl:512-515) and
 * mk_gate_nand (mk_gates.jl:7-12) over a batch in host memory, sharded over the devices */
int tfhe_b200_multi_gate_batch(tfhe_b200_multi* m, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                               int32_t* out, size_t count);
int tfhe_b200_multi_bootstrap_batch(tfhe_b200_multi* m, int32_t mu, const int32_t* x, int32_t* out, size_t count);
int tfhe_b200_multi_mk_nand_batch(tfhe_b200_multi* m, const int32_t* x, const int32_t* y, int32_t* out, size_t count);

#ifdef __cplusplus
}
#endif
#endif

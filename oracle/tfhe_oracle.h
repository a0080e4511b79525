/*
 * tfhe_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the gate-bootstrapping hot path of nucypher/TFHE.jl
 * (SURVEY.md §8a).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (tfhe.jl_b200/) never does.
 *
 * PARITY STATUS: the reference is pure Julia and Julia is not installed in this
 * image, so the reference itself cannot be run here.  The reference's own tests
 * hold no golden ciphertexts — only decrypted truth tables
 * (test/runtests.jl:8-57, 60-100; examples/tutorial.jl:77).  This oracle is pinned
 * against those truth tables and against the exact integer negacyclic convolution
 * (the mathematical ground truth the reference's FFT approximates,
 * polynomials.jl:138-140).  Ciphertext-level parity with TFHE.jl itself is
 * therefore "parity unpinned" (Julia's MersenneTwister stream cannot be
 * reproduced without Julia); see DESIGN.md §Oracle.
 *
 * Layouts (all int32 = Torus32, numeric-functions.jl:1):
 *   LWE ciphertext            [n+1]                a[0..n-1], b        (lwe.jl:21-29)
 *   BK (coefficient domain)   [n][l][k+1][k+1][N]  samples[r,j].a[c]   (tgsw.jl:28)
 *   KSK                       [N*k][t][base-1][n+1]                    (keyswitch.jl:36-38)
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t n;        /* lwe_size                 api.jl:7  */
    int32_t N;        /* tlwe_polynomial_degree   api.jl:10 */
    int32_t k;        /* tlwe_mask_size           api.jl:11 */
    int32_t l;        /* bs_decomp_length         api.jl:13 */
    int32_t bgbit;    /* bs_log2_base             api.jl:14 */
    int32_t t;        /* ks_decomp_length         api.jl:17 */
    int32_t basebit;  /* ks_log2_base             api.jl:18 */
    int32_t parties;  /* max_parties              api.jl:21 */
    double lwe_sigma; /* lwe_noise_stddev         api.jl:8  */
    double bs_sigma;  /* bs_noise_stddev          api.jl:15 */
    double ks_sigma;  /* ks_noise_stddev          api.jl:19 */
} orc_params;

/* gate opcodes, in the order of gates.jl */
enum {
    ORC_NAND = 0, ORC_OR = 1, ORC_AND = 2, ORC_XOR = 3, ORC_XNOR = 4, ORC_NOT = 5,
    ORC_CONSTANT = 6, ORC_NOR = 7, ORC_ANDNY = 8, ORC_ANDYN = 9, ORC_ORNY = 10,
    ORC_ORYN = 11, ORC_MUX = 12
};

/* polynomial-product route */
enum { ORC_ROUTE_EXACT = 0, /* O(N^2) integer negacyclic convolution (ground truth) */
       ORC_ROUTE_FFT = 1    /* the reference's folded complex-double FFT (polynomials.jl:106-144) */ };

typedef struct orc_rng orc_rng;
orc_rng* orc_rng_create(uint64_t seed);
void orc_rng_destroy(orc_rng*);
uint64_t orc_rng_u64(orc_rng*);
double orc_rng_normal(orc_rng*);

/* ---- L0/L1 primitives ---- */
int32_t orc_encode_message(int32_t mu, int32_t message_space);
int32_t orc_decode_message(int32_t phase, int32_t message_space);
int32_t orc_dtot32(double d);
void orc_mul_by_monomial(const int32_t* p, int64_t s, int32_t* out, int N);
void orc_reverse_polynomial(const int32_t* p, int32_t* out, int N);
void orc_polymul_exact(const int32_t* x, const int32_t* y, int32_t* out, int N);
void orc_polymul_fft(const int32_t* x, const int32_t* y, int32_t* out, int N);
void orc_forward_transform(const int32_t* p, double* out_re_im /* N/2 complex interleaved */, int N);
void orc_inverse_transform(const double* in_re_im, int32_t* out, int N);
int32_t orc_decomp_offset(int l, int bgbit);
void orc_decompose(const int32_t* p, int N, int l, int bgbit, int32_t* out /* [l][N] */);

/* ---- key generation / encrypt / decrypt (host-keep side of the reference; test fixtures) ---- */
void orc_keygen(const orc_params* P, uint64_t seed,
                int32_t* lwe_key /* [n] */, int32_t* tlwe_key /* [k][N] */,
                int32_t* bk /* [n][l][k+1][k+1][N] */,
                int32_t* ksk /* [N*k][t][base-1][n+1] */);
void orc_lwe_encrypt(orc_rng* rng, int32_t message, double alpha, const int32_t* key, int n,
                     int32_t* out /* [n+1] */);
int32_t orc_lwe_phase(const int32_t* ct, const int32_t* key, int n);

/* ---- hot path ---- */
typedef struct orc_ctx orc_ctx;
orc_ctx* orc_create(const orc_params* P, const int32_t* bk, const int32_t* ksk);
void orc_destroy(orc_ctx*);

/* tgsw_extern_mul (tgsw.jl:125-129): acc [k+1][N] (x) BK_i -> out [k+1][N] */
void orc_extern_mul(const orc_ctx* C, int bk_index, const int32_t* acc, int32_t* out, int route);
/* blind_rotate_and_extract pieces (bootstrap.jl:32-59); acc in/out [k+1][N] */
void orc_blind_rotate(const orc_ctx* C, int32_t* acc, const int32_t* bara, int route, int n_iter);
void orc_tlwe_extract(const int32_t* acc, int k, int N, int32_t* out /* [N*k+1] */);
/* bootstrap_wo_keyswitch (bootstrap.jl:69-82): x [n+1] -> out [N*k+1] */
void orc_bootstrap_wo_ks(const orc_ctx* C, int32_t mu, const int32_t* x, int32_t* out, int route);
/* keyswitch (keyswitch.jl:45-80): in [N*k+1] -> out [n+1] */
void orc_keyswitch(const orc_ctx* C, const int32_t* in, int32_t* out);
/* bootstrap (bootstrap.jl:92-95) */
void orc_bootstrap(const orc_ctx* C, int32_t mu, const int32_t* x, int32_t* out, int route);
/* gates.jl; x,y,z,out [count][n+1]; unused inputs may be NULL.  For ORC_CONSTANT, x[0] != 0 selects true. */
void orc_gate_batch(const orc_ctx* C, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                    int32_t* out, size_t count, int route, int nthreads);
/* gate prologue only (the linear combination before bootstrap): out [n+1] */
void orc_gate_prologue(int op, const int32_t* x, const int32_t* y, int n, int32_t* out);

/* ---- multi-key (mk_internals.jl, mk_api.jl, mk_gates.jl) ---- */
/* MK LWE ciphertext [parties*n + 1]: a[p][n] (Julia a[:,p] contiguous), then b.
 * MK BK (coefficient domain) [parties][n] samples (bk.key[j,i] -> [i][j]), each
 *   { x[l][p][N], y[l][p][N], c0[l][N], c1[l][N] }  = l*(2p+2) polys.
 * MK KSK: parties single-key KSKs back to back. */
void orc_mk_keygen(const orc_params* P, int parties, uint64_t seed,
                   int32_t* lwe_keys /* [p][n] */, int32_t* mk_bk, int32_t* mk_ksk /* [p][N][t][base-1][n+1] */,
                   int32_t* tlwe_keys_out /* [p][N] or NULL */);
size_t orc_mk_bk_words(const orc_params* P, int parties);
void orc_mk_encrypt(orc_rng* rng, const orc_params* P, int parties, const int32_t* lwe_keys, int message,
                    int32_t* out /* [p*n+1] */);
int32_t orc_mk_phase(const orc_params* P, int parties, const int32_t* lwe_keys, const int32_t* ct);

typedef struct orc_mk_ctx orc_mk_ctx;
orc_mk_ctx* orc_mk_create(const orc_params* P, int parties, const int32_t* mk_bk, const int32_t* mk_ksk);
void orc_mk_destroy(orc_mk_ctx*);
/* mk_tgsw_extern_mul (mk_internals.jl:348-391): acc [(p+1)][N] (a_1..a_p, b) */
void orc_mk_extern_mul(const orc_mk_ctx* C, int party, int bk_index, const int32_t* acc, int32_t* out, int route);
void orc_mk_bootstrap_wo_ks(const orc_mk_ctx* C, int32_t mu, const int32_t* x, int32_t* out /* [p*N+1] */, int route);
void orc_mk_keyswitch(const orc_mk_ctx* C, const int32_t* in /* [p*N+1] */, int32_t* out /* [p*n+1] */);
void orc_mk_nand_batch(const orc_mk_ctx* C, const int32_t* x, const int32_t* y, int32_t* out, size_t count,
                       int route, int nthreads);

#ifdef __cplusplus
}
#endif
#endif

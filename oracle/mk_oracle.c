/*
 * mk_oracle.c — CPU ORACLE, multi-key (MK-TFHE) part.  Test infrastructure, NOT product code
 * (see tfhe_oracle.h: "parity unpinned" at ciphertext level; pinned on the reference's MK NAND
 * test, test/runtests.jl:60-100, and on exact integer arithmetic).
 *
 * Restates src/mk_internals.jl, src/mk_api.jl, src/mk_gates.jl of nucypher/TFHE.jl.
 *
 * Layouts:
 *   MK LWE ciphertext [p*n + 1]: a[party][n] (Julia a[:,party] is contiguous, mk_internals.jl:9), then b.
 *   MK BK sample (coefficient domain), one per (party i, LWE index j) (bk.key[j,i], :453-455),
 *   stored [i][j]:  x[l][p][N] | y[l][p][N] | c0[l][N] | c1[l][N]      (mk_internals.jl:240-250)
 *   MK KSK: p single-key KSKs [p][N*k][t][base-1][n+1]                 (mk_api.jl:66-71)
 */
#include "tfhe_oracle_internal.h"

struct orc_mk_ctx {
    orc_params P;
    int parties;
    const int32_t* bk;
    const int32_t* ksk;
    cplx* bk_fft;      /* forward_transform.(samples), mk_internals.jl:457 */
};

static size_t mk_sample_polys(const orc_params* P, int p) { return (size_t)P->l * (2 * p + 2); }
size_t orc_mk_bk_words(const orc_params* P, int parties) {
    return (size_t)parties * P->n * mk_sample_polys(P, parties) * P->N;
}
static inline void poly_add(int32_t* dst, const int32_t* src, int N) {
    for (int m = 0; m < N; m++) dst[m] = (int32_t)((uint32_t)dst[m] + (uint32_t)src[m]);
}
static void poly_gauss(orc_rng* rng, double alpha, int32_t* out, int N) {
    for (int m = 0; m < N; m++) out[m] = orc_dtot32(orc_rng_normal(rng) * alpha);
}
static void poly_uniform(orc_rng* rng, int32_t* out, int N) { for (int m = 0; m < N; m++) out[m] = orc_rng_torus(rng); }

/* mk_api.jl:44-101 + mk_internals.jl:106-138 (SharedKey, PublicKey), :185-227 (RGSW.UniEnc),
 * :304-345 (RGSW.Expand), :426-460 (BootstrapKeyPart, MKBootstrapKey); keyswitch.jl:14-41 per party */
void orc_mk_keygen(const orc_params* P, int p, uint64_t seed, int32_t* lwe_keys, int32_t* mk_bk, int32_t* mk_ksk,
                   int32_t* tlwe_keys_out /* [p][N], may be NULL (test hook) */) {
    orc_rng* rng = orc_rng_create(seed);
    int n = P->n, N = P->N, l = P->l;
    size_t LN = (size_t)l * N;
    for (int i = 0; i < p * n; i++) lwe_keys[i] = orc_rng_bit(rng);              /* SecretKey, api.jl:96-99 */
    int32_t* shared_a = (int32_t*)malloc(sizeof(int32_t) * LN);                  /* mk_internals.jl:109 */
    for (int r = 0; r < l; r++) poly_uniform(rng, shared_a + (size_t)r * N, N);
    int32_t* tlwe_keys = (int32_t*)malloc(sizeof(int32_t) * (size_t)p * N);
    int32_t* pub_b = (int32_t*)malloc(sizeof(int32_t) * (size_t)p * LN);
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * N);
    /* uni-encryptions: per party, per LWE index: c0,c1,d0,d1,f0,f1 each [l][N] */
    size_t ue_words = 6 * LN;
    int32_t* ue = (int32_t*)malloc(sizeof(int32_t) * (size_t)p * n * ue_words);
    int32_t* rpoly = (int32_t*)malloc(sizeof(int32_t) * N);
    size_t ksk_words = (size_t)N * P->k * P->t * ((1 << P->basebit) - 1) * (n + 1);

    for (int i = 0; i < p; i++) {
        int32_t* S = tlwe_keys + (size_t)i * N;
        for (int m = 0; m < N; m++) S[m] = orc_rng_bit(rng);                     /* mk_api.jl:69 */
        for (int r = 0; r < l; r++) {                                            /* mk_internals.jl:132-135 */
            int32_t* b = pub_b + (size_t)i * LN + (size_t)r * N;
            orc_polymul_fft(S, shared_a + (size_t)r * N, b, N);
            poly_gauss(rng, P->bs_sigma, tmp, N); poly_add(b, tmp, N);
        }
        for (int j = 0; j < n; j++) {                                            /* mk_internals.jl:433-435 */
            int32_t msg = lwe_keys[(size_t)i * n + j];
            int32_t* u = ue + ((size_t)i * n + j) * ue_words;
            int32_t *c0 = u, *c1 = u + LN, *d0 = u + 2 * LN, *d1 = u + 3 * LN, *f0 = u + 4 * LN, *f1 = u + 5 * LN;
            for (int m = 0; m < N; m++) rpoly[m] = orc_rng_bit(rng);             /* :195 */
            for (int r = 0; r < l; r++) {
                uint32_t g = 1u << (32 - (r + 1) * P->bgbit);
                size_t o = (size_t)r * N;
                poly_uniform(rng, c1 + o, N);                                    /* :198 */
                poly_gauss(rng, P->bs_sigma, c0 + o, N);                         /* :200-204 */
                orc_polymul_fft(S, c1 + o, tmp, N); poly_add(c0 + o, tmp, N);
                c0[o] = (int32_t)((uint32_t)c0[o] + (uint32_t)msg * g);
                poly_gauss(rng, P->bs_sigma, d1 + o, N);                         /* :207-211 */
                orc_polymul_fft(rpoly, shared_a + o, tmp, N); poly_add(d1 + o, tmp, N);
                d1[o] = (int32_t)((uint32_t)d1[o] + (uint32_t)msg * g);
                poly_gauss(rng, P->bs_sigma, d0 + o, N);                         /* :212-215 */
                orc_polymul_fft(rpoly, pub_b + (size_t)i * LN + o, tmp, N); poly_add(d0 + o, tmp, N);
                poly_uniform(rng, f1 + o, N);                                    /* :218 */
                poly_gauss(rng, P->bs_sigma, f0 + o, N);                         /* :220-224 */
                orc_polymul_fft(S, f1 + o, tmp, N); poly_add(f0 + o, tmp, N);
                for (int m = 0; m < N; m++) f0[o + m] = (int32_t)((uint32_t)f0[o + m] + (uint32_t)rpoly[m] * g);
            }
        }
        /* per-party keyswitch key (mk_api.jl:70-72 -> keyswitch.jl:14-41) */
        {
            int t = P->t, base = 1 << P->basebit, Nk = N * P->k;
            int32_t* ksk = mk_ksk + (size_t)i * ksk_words;
            const int32_t* sk = lwe_keys + (size_t)i * n;
            size_t cnt = (size_t)Nk * t * (base - 1);
            double* noise = (double*)malloc(sizeof(double) * cnt);
            double sum = 0;
            for (size_t q = 0; q < cnt; q++) { noise[q] = orc_rng_normal(rng) * P->ks_sigma; sum += noise[q]; }
            for (size_t q = 0; q < cnt; q++) noise[q] -= sum / (double)cnt;
            for (int a = 0; a < Nk; a++)
                for (int jj = 0; jj < t; jj++)
                    for (int h = 1; h < base; h++) {
                        size_t q = ((size_t)a * t + jj) * (base - 1) + (h - 1);
                        int32_t* row = ksk + q * (n + 1);
                        uint32_t m = ((uint32_t)S[a] * (uint32_t)h) << (32 - (jj + 1) * P->basebit);
                        uint32_t dot = 0;
                        for (int c = 0; c < n; c++) { row[c] = orc_rng_torus(rng); dot += (uint32_t)row[c] * (uint32_t)sk[c]; }
                        row[n] = (int32_t)(m + (uint32_t)orc_dtot32(noise[q]) + dot);
                    }
            free(noise);
        }
    }
    /* RGSW.Expand (mk_internals.jl:304-345) */
    size_t spolys = mk_sample_polys(P, p);
    int32_t* diff = (int32_t*)malloc(sizeof(int32_t) * N);
    int32_t* dec = (int32_t*)malloc(sizeof(int32_t) * LN);
    for (int i = 0; i < p; i++)
        for (int j = 0; j < n; j++) {
            const int32_t* u = ue + ((size_t)i * n + j) * ue_words;
            const int32_t *c0 = u, *c1 = u + LN, *d0 = u + 2 * LN, *d1 = u + 3 * LN, *f0 = u + 4 * LN, *f1 = u + 5 * LN;
            int32_t* s = mk_bk + ((size_t)i * n + j) * spolys * N;
            int32_t* x = s;
            int32_t* y = s + (size_t)l * p * N;
            memcpy(s + (size_t)2 * l * p * N, c0, sizeof(int32_t) * LN);
            memcpy(s + (size_t)2 * l * p * N + LN, c1, sizeof(int32_t) * LN);
            for (int jj = 0; jj < l; jj++)
                for (int ii = 0; ii < p; ii++) {
                    int32_t* xo = x + ((size_t)jj * p + ii) * N;
                    int32_t* yo = y + ((size_t)jj * p + ii) * N;
                    memcpy(xo, d0 + (size_t)jj * N, sizeof(int32_t) * N);        /* :327 */
                    if (ii == i) { memcpy(yo, d1 + (size_t)jj * N, sizeof(int32_t) * N); continue; }   /* :336 */
                    const int32_t* bi = pub_b + (size_t)ii * LN + (size_t)jj * N;
                    const int32_t* bp = pub_b + (size_t)i * LN + (size_t)jj * N;
                    for (int m = 0; m < N; m++) diff[m] = (int32_t)((uint32_t)bi[m] - (uint32_t)bp[m]);
                    orc_decompose(diff, N, l, P->bgbit, dec);                    /* :321 */
                    memset(yo, 0, sizeof(int32_t) * N);
                    for (int r = 0; r < l; r++) {
                        orc_polymul_fft(dec + (size_t)r * N, f0 + (size_t)r * N, tmp, N); poly_add(xo, tmp, N);  /* :330 */
                        orc_polymul_fft(dec + (size_t)r * N, f1 + (size_t)r * N, tmp, N); poly_add(yo, tmp, N);  /* :338 */
                    }
                }
        }
    if (tlwe_keys_out) memcpy(tlwe_keys_out, tlwe_keys, sizeof(int32_t) * (size_t)p * N);
    free(diff); free(dec); free(rpoly); free(ue); free(tmp); free(pub_b); free(tlwe_keys); free(shared_a);
    orc_rng_destroy(rng);
}

/* mk_api.jl:110-126 */
void orc_mk_encrypt(orc_rng* rng, const orc_params* P, int p, const int32_t* lwe_keys, int message, int32_t* out) {
    int n = P->n;
    uint32_t dot = 0;
    for (int i = 0; i < p * n; i++) { out[i] = orc_rng_torus(rng); dot += (uint32_t)out[i] * (uint32_t)lwe_keys[i]; }
    int32_t mu = orc_encode_message(message ? 1 : -1, 8);
    out[p * n] = (int32_t)((uint32_t)mu + (uint32_t)orc_dtot32(orc_rng_normal(rng) * P->lwe_sigma) + dot);
}
/* mk_internals.jl:29-35 */
int32_t orc_mk_phase(const orc_params* P, int p, const int32_t* lwe_keys, const int32_t* ct) {
    uint32_t dot = 0;
    for (int i = 0; i < p * P->n; i++) dot += (uint32_t)ct[i] * (uint32_t)lwe_keys[i];
    return (int32_t)((uint32_t)ct[p * P->n] - dot);
}

orc_mk_ctx* orc_mk_create(const orc_params* P, int parties, const int32_t* mk_bk, const int32_t* mk_ksk) {
    orc_mk_ctx* C = (orc_mk_ctx*)calloc(1, sizeof(orc_mk_ctx));
    C->P = *P; C->parties = parties; C->bk = mk_bk; C->ksk = mk_ksk;
    size_t polys = (size_t)parties * P->n * mk_sample_polys(P, parties);
    int N = P->N;
    C->bk_fft = (cplx*)malloc(sizeof(cplx) * polys * (N / 2));
    orc_warm_plan(N);
    #pragma omp parallel for schedule(static)
    for (size_t q = 0; q < polys; q++) orc_forward_cplx(mk_bk + q * N, C->bk_fft + q * (N / 2), N);
    return C;
}
void orc_mk_destroy(orc_mk_ctx* C) { if (C) { free(C->bk_fft); free(C); } }

/* mk_internals.jl:348-391.  acc/out: [(p+1)][N] = a_1..a_p, b. */
void orc_mk_extern_mul(const orc_mk_ctx* C, int party, int j, const int32_t* acc, int32_t* out, int route) {
    int N = C->P.N, l = C->P.l, p = C->parties, n2 = N / 2;
    size_t spolys = mk_sample_polys(&C->P, p);
    size_t sidx = ((size_t)party * C->P.n + j) * spolys;      /* first poly of this sample */
    /* poly index helpers inside the sample */
    #define XI(r, ii) ((size_t)(r) * p + (ii))
    #define YI(r, ii) ((size_t)l * p + (size_t)(r) * p + (ii))
    #define C0I(r) ((size_t)2 * l * p + (r))
    #define C1I(r) ((size_t)2 * l * p + l + (r))
    int32_t* dec = (int32_t*)malloc(sizeof(int32_t) * (size_t)(p + 1) * l * N);   /* [(p+1)][l][N] */
    for (int q = 0; q <= p; q++) orc_decompose(acc + (size_t)q * N, N, l, C->P.bgbit, dec + (size_t)q * l * N);
    memset(out, 0, sizeof(int32_t) * (size_t)(p + 1) * N);
    int32_t* prod = (int32_t*)malloc(sizeof(int32_t) * N);
    cplx* tr = NULL; cplx* buf = NULL;
    if (route == ORC_ROUTE_FFT) {
        tr = (cplx*)malloc(sizeof(cplx) * (size_t)(p + 1) * l * n2);              /* :368-369 */
        buf = (cplx*)malloc(sizeof(cplx) * n2);
        for (int q = 0; q < (p + 1) * l; q++) orc_forward_cplx(dec + (size_t)q * N, tr + (size_t)q * n2, N);
    }
    /* one product dec[q][r] (*) sample_poly[si], added into out[o]; the reference inverse-transforms
     * every product separately and sums the integers (:359-366, :373-387) */
    #define MULADD(q, r, si, o) do { \
        if (route == ORC_ROUTE_EXACT) orc_polymul_exact(dec + ((size_t)(q) * l + (r)) * N, C->bk + (sidx + (si)) * N, prod, N); \
        else { const cplx* A = tr + ((size_t)(q) * l + (r)) * n2; const cplx* B = C->bk_fft + (sidx + (si)) * n2; \
               for (int f = 0; f < n2; f++) { buf[f] = A[f] * B[f]; } \
               orc_inverse_cplx(buf, prod, N); } \
        poly_add(out + (size_t)(o) * N, prod, N); } while (0)
    for (int ii = 0; ii < p; ii++) {
        if (ii == party) {
            for (int r = 0; r < l; r++) for (int jj = 0; jj < p; jj++) MULADD(jj, r, YI(r, jj), ii);   /* :375-376 */
            for (int r = 0; r < l; r++) MULADD(p, r, C1I(r), ii);                                       /* :377-378 */
        } else {
            for (int r = 0; r < l; r++) MULADD(ii, r, YI(r, party), ii);                                /* :379-380 */
        }
    }
    for (int r = 0; r < l; r++) for (int ii = 0; ii < p; ii++) MULADD(ii, r, XI(r, ii), p);             /* :384-385 */
    for (int r = 0; r < l; r++) MULADD(p, r, C0I(r), p);                                                /* :386-387 */
    #undef MULADD
    #undef XI
    #undef YI
    #undef C0I
    #undef C1I
    free(dec); free(prod); free(tr); free(buf);
}

/* mk_internals.jl:498-509, 488-495, 473-485, 464-470, 88-95 */
void orc_mk_bootstrap_wo_ks(const orc_mk_ctx* C, int32_t mu, const int32_t* x, int32_t* out, int route) {
    int n = C->P.n, N = C->P.N, p = C->parties;
    size_t sz = (size_t)(p + 1) * N;
    int32_t* acc = (int32_t*)calloc(sz, sizeof(int32_t));
    int32_t* temp = (int32_t*)malloc(sizeof(int32_t) * sz * 2);
    int32_t* prod = temp + sz;
    int32_t* tv = (int32_t*)malloc(sizeof(int32_t) * N);
    int32_t barb = orc_decode_message(x[p * n], 2 * N);                          /* :502 */
    for (int m = 0; m < N; m++) tv[m] = mu;                                      /* :506 */
    orc_mul_by_monomial(tv, -(int64_t)barb, acc + (size_t)p * N, N);             /* :491-492 */
    for (int i = 0; i < p; i++)                                                  /* :475 parties outer */
        for (int j = 0; j < n; j++) {                                            /* :476 */
            int32_t bara = orc_decode_message(x[(size_t)i * n + j], 2 * N);      /* :503 */
            if (bara == 0) continue;                                             /* :478 */
            for (int q = 0; q <= p; q++) {                                       /* :468 */
                orc_mul_by_monomial(acc + (size_t)q * N, bara, temp + (size_t)q * N, N);
                for (int m = 0; m < N; m++)
                    temp[(size_t)q * N + m] = (int32_t)((uint32_t)temp[(size_t)q * N + m] - (uint32_t)acc[(size_t)q * N + m]);
            }
            orc_mk_extern_mul(C, i, j, temp, prod, route);                       /* :469 */
            for (size_t m = 0; m < sz; m++) acc[m] = (int32_t)((uint32_t)acc[m] + (uint32_t)prod[m]);
        }
    for (int q = 0; q < p; q++) orc_reverse_polynomial(acc + (size_t)q * N, out + (size_t)q * N, N);   /* :91 */
    out[(size_t)p * N] = acc[(size_t)p * N];                                     /* :92 */
    free(acc); free(temp); free(tv);
}
/* mk_internals.jl:397-411 */
void orc_mk_keyswitch(const orc_mk_ctx* C, const int32_t* in, int32_t* out) {
    int n = C->P.n, N = C->P.N, p = C->parties, Nk = N * C->P.k;
    size_t ksk_words = (size_t)Nk * C->P.t * ((1 << C->P.basebit) - 1) * (n + 1);
    int32_t* tin = (int32_t*)malloc(sizeof(int32_t) * (Nk + 1));
    int32_t* tout = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
    uint32_t b = (uint32_t)in[(size_t)p * N];
    for (int q = 0; q < p; q++) {
        memcpy(tin, in + (size_t)q * N, sizeof(int32_t) * Nk);
        tin[Nk] = 0;                                                             /* :400 */
        orc_keyswitch_raw(C->ksk + (size_t)q * ksk_words, Nk, n, C->P.t, C->P.basebit, tin, tout);
        memcpy(out + (size_t)q * n, tout, sizeof(int32_t) * n);
        b += (uint32_t)tout[n];                                                  /* :409 */
    }
    out[(size_t)p * n] = (int32_t)b;
    free(tin); free(tout);
}
/* mk_gates.jl:7-12 + mk_internals.jl:512-515 */
void orc_mk_nand_batch(const orc_mk_ctx* C, const int32_t* x, const int32_t* y, int32_t* out, size_t count,
                       int route, int nthreads) {
    int n = C->P.n, N = C->P.N, p = C->parties;
    size_t w = (size_t)p * n + 1;
    if (nthreads < 1) nthreads = 1;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (size_t g = 0; g < count; g++) {
        int32_t* lin = (int32_t*)malloc(sizeof(int32_t) * w);
        int32_t* u = (int32_t*)malloc(sizeof(int32_t) * ((size_t)p * N + 1));
        for (size_t c = 0; c < w; c++) lin[c] = (int32_t)(0u - (uint32_t)x[g * w + c] - (uint32_t)y[g * w + c]);
        lin[w - 1] = (int32_t)((uint32_t)lin[w - 1] + (uint32_t)orc_encode_message(1, 8));
        orc_mk_bootstrap_wo_ks(C, orc_encode_message(1, 8), lin, u, route);
        orc_mk_keyswitch(C, u, out + g * w);
        free(lin); free(u);
    }
}

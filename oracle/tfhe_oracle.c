/*
 * tfhe_oracle.c — CPU ORACLE, single-key part (test infrastructure, NOT product code).
 * See tfhe_oracle.h for scope and parity status ("parity unpinned" at ciphertext level;
 * pinned on the reference's truth-table tests and on exact integer arithmetic).
 *
 * Every function cites the reference file:line (under /root/reference/src) it restates.
 * Integer arithmetic is two's-complement int32 with wrap-around (done in uint32_t to stay
 * defined in C); shifts on Torus32 are arithmetic, as in Julia.
 */
#include "tfhe_oracle_internal.h"

/* ------------------------------------------------------------------ RNG (fixtures only) */
/* xoshiro256** seeded by splitmix64; Gaussian by Box-Muller.  The reference uses Julia's
 * MersenneTwister (test/runtests.jl:27), whose stream is not reproducible here. */
struct orc_rng { uint64_t s[4]; int have_spare; double spare; };

static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
orc_rng* orc_rng_create(uint64_t seed) {
    orc_rng* r = (orc_rng*)calloc(1, sizeof(orc_rng));
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed);
    return r;
}
void orc_rng_destroy(orc_rng* r) { free(r); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
uint64_t orc_rng_u64(orc_rng* r) {
    uint64_t* s = r->s;
    uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return result;
}
static inline double rng_unit(orc_rng* r) { return ((orc_rng_u64(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
double orc_rng_normal(orc_rng* r) {
    if (r->have_spare) { r->have_spare = 0; return r->spare; }
    double u = rng_unit(r), v = rng_unit(r);
    double m = sqrt(-2.0 * log(u));
    r->spare = m * sin(2.0 * M_PI * v); r->have_spare = 1;
    return m * cos(2.0 * M_PI * v);
}
int32_t orc_rng_torus(orc_rng* r) { return (int32_t)(uint32_t)(orc_rng_u64(r) >> 32); } /* numeric-functions.jl:9-11 */
int32_t orc_rng_bit(orc_rng* r) { return (int32_t)(orc_rng_u64(r) >> 63); }             /* numeric-functions.jl:4-6 */

/* ------------------------------------------------------------------ L0: torus scalars */
/* numeric-functions.jl:42-45 */
int32_t orc_encode_message(int32_t mu, int32_t message_space) {
    int log2_ms = __builtin_ctz((unsigned)message_space);
    return (int32_t)((uint32_t)mu << (32 - log2_ms));
}
/* numeric-functions.jl:31-34 — also the modulus switch of bootstrap.jl:74-75 */
int32_t orc_decode_message(int32_t phase, int32_t message_space) {
    int log2_ms = __builtin_ctz((unsigned)message_space);
    int32_t shifted = (int32_t)((uint32_t)phase + (1u << (32 - log2_ms - 1)));
    return shifted >> (32 - log2_ms);
}
/* numeric-functions.jl:51-53 (trunc toward zero) */
int32_t orc_dtot32(double d) { return (int32_t)(d * 4294967296.0); }

/* ------------------------------------------------------------------ L1: polynomials */
/* DarkIntegers mul_by_monomial: X^s * p mod X^N+1, s any integer */
void orc_mul_by_monomial(const int32_t* p, int64_t s, int32_t* out, int N) {
    int64_t twoN = 2 * (int64_t)N;
    s %= twoN; if (s < 0) s += twoN;
    int neg = 0;
    if (s >= N) { neg = 1; s -= N; }
    for (int m = 0; m < N; m++) {
        uint32_t v;
        if (m >= s) v = (uint32_t)p[m - s];
        else v = 0u - (uint32_t)p[m - s + N];
        out[m] = (int32_t)(neg ? 0u - v : v);
    }
}
/* polynomials.jl:32-35: reverse the coefficient array, then multiply by X^(N+1) */
void orc_reverse_polynomial(const int32_t* p, int32_t* out, int N) {
    int32_t* tmp = (int32_t*)malloc(sizeof(int32_t) * N);
    for (int i = 0; i < N; i++) tmp[i] = p[N - 1 - i];
    orc_mul_by_monomial(tmp, N + 1, out, N);
    free(tmp);
}
/* Ground truth: exact negacyclic convolution mod 2^32.  Wrap-around
 * uint32 arithmetic is exact mod 2^32, which is all the torus keeps. */
void orc_polymul_exact(const int32_t* x, const int32_t* y, int32_t* out, int N) {
    uint32_t* acc = (uint32_t*)calloc(N, sizeof(uint32_t));
    for (int i = 0; i < N; i++) {
        uint32_t xi = (uint32_t)x[i];
        if (!xi) continue;
        for (int j = 0; j < N - i; j++) acc[i + j] += xi * (uint32_t)y[j];
        for (int j = N - i; j < N; j++) acc[i + j - N] -= xi * (uint32_t)y[j];
    }
    for (int i = 0; i < N; i++) out[i] = (int32_t)acc[i];
    free(acc);
}

/* --- the reference's transform (polynomials.jl:44-132): folded N/2-point complex FFT --- */
typedef struct { int N; cplx* twist; /* exp(-i*pi*j/N), polynomials.jl:53,71 */
                 cplx* w;     /* exp(-2*pi*i*j/(N/2)) */ int* rev; } fft_plan;
#define MAX_PLANS 8
static fft_plan g_plans[MAX_PLANS];
static int g_nplans = 0;

static const fft_plan* get_plan(int N) {
    const fft_plan* found = NULL;
    #pragma omp critical(orc_plan)
    {
        for (int i = 0; i < g_nplans; i++) if (g_plans[i].N == N) found = &g_plans[i];
        if (!found && g_nplans < MAX_PLANS) {
            fft_plan* p = &g_plans[g_nplans];
            int n = N / 2, lg = 0;
            while ((1 << lg) < n) lg++;
            p->N = N;
            p->twist = (cplx*)malloc(sizeof(cplx) * n);
            p->w = (cplx*)malloc(sizeof(cplx) * n);
            p->rev = (int*)malloc(sizeof(int) * n);
            for (int j = 0; j < n; j++) {
                double a = -M_PI * j / N;
                p->twist[j] = cos(a) + I * sin(a);
                double b = -2.0 * M_PI * j / n;
                p->w[j] = cos(b) + I * sin(b);
                int r = 0;
                for (int b2 = 0; b2 < lg; b2++) if (j & (1 << b2)) r |= 1 << (lg - 1 - b2);
                p->rev[j] = r;
            }
            g_nplans++;
            found = p;
        }
    }
    return found;
}

void orc_warm_plan(int N) { (void)get_plan(N); }

/* in-place radix-2 DIT FFT of length n = N/2; sign = -1 forward (plan_fft), +1 inverse
 * (un-normalised; the caller divides, as plan_ifft does, polynomials.jl:74) */
static void fft_inplace(cplx* a, const fft_plan* p, int sign) {
    int n = p->N / 2;
    for (int i = 0; i < n; i++) { int r = p->rev[i]; if (r > i) { cplx t = a[i]; a[i] = a[r]; a[r] = t; } }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int j = 0; j < half; j++) {
                cplx w = p->w[j * step];
                if (sign > 0) w = conj(w);
                cplx u = a[i + j], v = a[i + j + half] * w;
                a[i + j] = u + v; a[i + j + half] = u - v;
            }
    }
}
/* polynomials.jl:106-112 */
void orc_forward_cplx(const int32_t* c, cplx* out, int N) {
    const fft_plan* p = get_plan(N);
    int n = N / 2;
    for (int j = 0; j < n; j++) out[j] = ((double)c[j] - I * (double)c[j + n]) * p->twist[j];
    fft_inplace(out, p, -1);
}
/* polynomials.jl:115-116: round to nearest (ties to even, Julia round(Int64, x)), keep the low 32 bits */
static inline int32_t to_int32(double x) { return (int32_t)(uint32_t)(uint64_t)llrint(x); }
/* polynomials.jl:119-132 (destroys `in`) */
void orc_inverse_cplx(cplx* in, int32_t* out, int N) {
    const fft_plan* p = get_plan(N);
    int n = N / 2;
    fft_inplace(in, p, +1);
    double inv = 1.0 / n;
    for (int j = 0; j < n; j++) {
        cplx u = conj(in[j] * inv) * p->twist[j];
        out[j] = to_int32(creal(u));
        out[j + n] = to_int32(cimag(u));
    }
}
void orc_forward_transform(const int32_t* p, double* out, int N) { orc_forward_cplx(p, (cplx*)out, N); }
void orc_inverse_transform(const double* in, int32_t* out, int N) {
    int n = N / 2;
    cplx* tmp = (cplx*)malloc(sizeof(cplx) * n);
    memcpy(tmp, in, sizeof(cplx) * n);
    orc_inverse_cplx(tmp, out, N);
    free(tmp);
}
/* polynomials.jl:142-144 */
void orc_polymul_fft(const int32_t* x, const int32_t* y, int32_t* out, int N) {
    int n = N / 2;
    cplx* a = (cplx*)malloc(sizeof(cplx) * n * 2);
    cplx* b = a + n;
    orc_forward_cplx(x, a, N);
    orc_forward_cplx(y, b, N);
    for (int j = 0; j < n; j++) a[j] *= b[j];
    orc_inverse_cplx(a, out, N);
    free(a);
}

/* ------------------------------------------------------------------ L2: TGSW pieces */
/* tgsw.jl:14,18: offset = (Bg/2) * sum_{r=1..l} 2^(32 - r*bgbit), wrapped to int32 */
int32_t orc_decomp_offset(int l, int bgbit) {
    uint32_t sum = 0;
    for (int r = 1; r <= l; r++) sum += 1u << (32 - r * bgbit);
    return (int32_t)(sum * (1u << (bgbit - 1)));
}
/* tgsw.jl:99-117 */
void orc_decompose(const int32_t* p, int N, int l, int bgbit, int32_t* out) {
    int32_t mask = (1 << bgbit) - 1, half = 1 << (bgbit - 1);
    uint32_t offset = (uint32_t)orc_decomp_offset(l, bgbit);
    for (int r = 1; r <= l; r++)
        for (int m = 0; m < N; m++) {
            int32_t v = (int32_t)((uint32_t)p[m] + offset);
            out[(size_t)(r - 1) * N + m] = ((v >> (32 - r * bgbit)) & mask) - half;
        }
}

/* ------------------------------------------------------------------ encrypt / decrypt / keygen */
/* lwe.jl:38-43 (rand_gaussian_torus32: numeric-functions.jl:20-23) */
void orc_lwe_encrypt(orc_rng* rng, int32_t message, double alpha, const int32_t* key, int n, int32_t* out) {
    uint32_t dot = 0;
    for (int i = 0; i < n; i++) { out[i] = orc_rng_torus(rng); dot += (uint32_t)out[i] * (uint32_t)key[i]; }
    out[n] = (int32_t)((uint32_t)message + (uint32_t)orc_dtot32(orc_rng_normal(rng) * alpha) + dot);
}
/* lwe.jl:59 */
int32_t orc_lwe_phase(const int32_t* ct, const int32_t* key, int n) {
    uint32_t dot = 0;
    for (int i = 0; i < n; i++) dot += (uint32_t)ct[i] * (uint32_t)key[i];
    return (int32_t)((uint32_t)ct[n] - dot);
}
/* tlwe.jl:63-73: (a_1..a_k uniform, b = e + sum_j S_j (*) a_j); out [k+1][N] */
void orc_tlwe_encrypt_zero(orc_rng* rng, double alpha, const int32_t* tlwe_key, int k, int N, int32_t* out) {
    int32_t* prod = (int32_t*)malloc(sizeof(int32_t) * N);
    int32_t* b = out + (size_t)k * N;
    for (int j = 0; j < k; j++) for (int m = 0; m < N; m++) out[(size_t)j * N + m] = orc_rng_torus(rng);
    for (int m = 0; m < N; m++) b[m] = orc_dtot32(orc_rng_normal(rng) * alpha);
    for (int j = 0; j < k; j++) {
        orc_polymul_fft(tlwe_key + (size_t)j * N, out + (size_t)j * N, prod, N);
        for (int m = 0; m < N; m++) b[m] = (int32_t)((uint32_t)b[m] + (uint32_t)prod[m]);
    }
    free(prod);
}
/* api.jl:96-99,116-126; bootstrap.jl:6-15 (tgsw.jl:52-88); keyswitch.jl:14-41 */
void orc_keygen(const orc_params* P, uint64_t seed, int32_t* lwe_key, int32_t* tlwe_key, int32_t* bk, int32_t* ksk) {
    orc_rng* rng = orc_rng_create(seed);
    int n = P->n, N = P->N, k = P->k, l = P->l;
    for (int i = 0; i < n; i++) lwe_key[i] = orc_rng_bit(rng);               /* lwe.jl:10-12 */
    for (int i = 0; i < k * N; i++) tlwe_key[i] = orc_rng_bit(rng);           /* tlwe.jl:15-20 */
    /* BK_i = TGSW(s_i): l x (k+1) zero-encryptions + s_i * gadget on the diagonal (tgsw.jl:62-69) */
    size_t sample = (size_t)(k + 1) * N;
    for (int i = 0; i < n; i++)
        for (int r = 0; r < l; r++)
            for (int j = 0; j <= k; j++) {
                int32_t* s = bk + (((size_t)i * l + r) * (k + 1) + j) * sample;
                orc_tlwe_encrypt_zero(rng, P->bs_sigma, tlwe_key, k, N, s);
                uint32_t g = 1u << (32 - (r + 1) * P->bgbit);                 /* tgsw.jl:14 */
                s[(size_t)j * N] = (int32_t)((uint32_t)s[(size_t)j * N] + (uint32_t)lwe_key[i] * g); /* coeff 0 of component j */
            }
    /* KSK (keyswitch.jl:27-38): centred noise, then LWE_s((s'_i*h) << (32 - j*basebit)) */
    int t = P->t, base = 1 << P->basebit, Nk = N * k;
    size_t cnt = (size_t)Nk * t * (base - 1);
    double* noise = (double*)malloc(sizeof(double) * cnt);
    double sum = 0;
    for (size_t q = 0; q < cnt; q++) { noise[q] = orc_rng_normal(rng) * P->ks_sigma; sum += noise[q]; }
    double mean = sum / (double)cnt;
    for (size_t q = 0; q < cnt; q++) noise[q] -= mean;
    for (int i = 0; i < Nk; i++)
        for (int j = 0; j < t; j++)
            for (int h = 1; h < base; h++) {
                size_t q = ((size_t)i * t + j) * (base - 1) + (h - 1);
                int32_t* row = ksk + q * (n + 1);
                uint32_t msg = ((uint32_t)tlwe_key[i] * (uint32_t)h) << (32 - (j + 1) * P->basebit);
                uint32_t dot = 0;
                for (int c = 0; c < n; c++) { row[c] = orc_rng_torus(rng); dot += (uint32_t)row[c] * (uint32_t)lwe_key[c]; }
                row[n] = (int32_t)(msg + (uint32_t)orc_dtot32(noise[q]) + dot);   /* lwe.jl:49-55 */
            }
    free(noise);
    orc_rng_destroy(rng);
}

/* ------------------------------------------------------------------ L3: bootstrapping engine */
orc_ctx* orc_create(const orc_params* P, const int32_t* bk, const int32_t* ksk) {
    orc_ctx* C = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    C->P = *P; C->bk = bk; C->ksk = ksk;
    int n = P->n, N = P->N, k = P->k, l = P->l;
    /* bootstrap.jl:12: the reference keeps the TRANSFORMED key */
    size_t polys = (size_t)n * l * (k + 1) * (k + 1);
    C->bk_fft = (cplx*)malloc(sizeof(cplx) * polys * (N / 2));
    orc_warm_plan(N);
    #pragma omp parallel for schedule(static)
    for (size_t q = 0; q < polys; q++) orc_forward_cplx(bk + q * N, C->bk_fft + q * (N / 2), N);
    return C;
}
void orc_destroy(orc_ctx* C) { if (C) { free(C->bk_fft); free(C); } }

/* tgsw.jl:125-129 */
void orc_extern_mul(const orc_ctx* C, int i, const int32_t* acc, int32_t* out, int route) {
    int N = C->P.N, k = C->P.k, l = C->P.l, n2 = N / 2;
    int32_t* dec = (int32_t*)malloc(sizeof(int32_t) * (size_t)l * N);
    if (route == ORC_ROUTE_EXACT) {
        int32_t* prod = (int32_t*)malloc(sizeof(int32_t) * N);
        memset(out, 0, sizeof(int32_t) * (size_t)(k + 1) * N);
        for (int j = 0; j <= k; j++) {
            orc_decompose(acc + (size_t)j * N, N, l, C->P.bgbit, dec);
            for (int r = 0; r < l; r++)
                for (int c = 0; c <= k; c++) {
                    const int32_t* b = C->bk + ((((size_t)i * l + r) * (k + 1) + j) * (k + 1) + c) * N;
                    orc_polymul_exact(dec + (size_t)r * N, b, prod, N);
                    for (int m = 0; m < N; m++)
                        out[(size_t)c * N + m] = (int32_t)((uint32_t)out[(size_t)c * N + m] + (uint32_t)prod[m]);
                }
        }
        free(prod);
    } else {
        /* sum in the transform domain, one inverse per output polynomial (tgsw.jl:128) */
        cplx* sum = (cplx*)calloc((size_t)(k + 1) * n2, sizeof(cplx));
        cplx* d = (cplx*)malloc(sizeof(cplx) * n2);
        for (int j = 0; j <= k; j++) {
            orc_decompose(acc + (size_t)j * N, N, l, C->P.bgbit, dec);
            for (int r = 0; r < l; r++) {
                orc_forward_cplx(dec + (size_t)r * N, d, N);
                for (int c = 0; c <= k; c++) {
                    const cplx* b = C->bk_fft + ((((size_t)i * l + r) * (k + 1) + j) * (k + 1) + c) * n2;
                    cplx* s = sum + (size_t)c * n2;
                    for (int f = 0; f < n2; f++) s[f] += d[f] * b[f];
                }
            }
        }
        for (int c = 0; c <= k; c++) orc_inverse_cplx(sum + (size_t)c * n2, out + (size_t)c * N, N);
        free(sum); free(d);
    }
    free(dec);
}

/* bootstrap.jl:19-23 (mux_rotate) inside bootstrap.jl:32-39 (blind_rotate) */
void orc_blind_rotate(const orc_ctx* C, int32_t* acc, const int32_t* bara, int route, int n_iter) {
    int N = C->P.N, k = C->P.k;
    size_t sz = (size_t)(k + 1) * N;
    int32_t* temp = (int32_t*)malloc(sizeof(int32_t) * sz * 2);
    int32_t* prod = temp + sz;
    for (int i = 0; i < n_iter; i++) {
        if (bara[i] == 0) continue;                                             /* bootstrap.jl:34 */
        for (int c = 0; c <= k; c++) {
            orc_mul_by_monomial(acc + (size_t)c * N, bara[i], temp + (size_t)c * N, N);   /* tlwe.jl:92-93 */
            for (int m = 0; m < N; m++)
                temp[(size_t)c * N + m] = (int32_t)((uint32_t)temp[(size_t)c * N + m] - (uint32_t)acc[(size_t)c * N + m]);
        }
        orc_extern_mul(C, i, temp, prod, route);
        for (size_t m = 0; m < sz; m++) acc[m] = (int32_t)((uint32_t)acc[m] + (uint32_t)prod[m]);
    }
    free(temp);
}
/* tlwe.jl:55-59 */
void orc_tlwe_extract(const int32_t* acc, int k, int N, int32_t* out) {
    for (int j = 0; j < k; j++) orc_reverse_polynomial(acc + (size_t)j * N, out + (size_t)j * N, N);
    out[(size_t)k * N] = acc[(size_t)k * N];
}
/* bootstrap.jl:69-82 + 50-59 */
void orc_bootstrap_wo_ks(const orc_ctx* C, int32_t mu, const int32_t* x, int32_t* out, int route) {
    int n = C->P.n, N = C->P.N, k = C->P.k;
    int32_t* bara = (int32_t*)malloc(sizeof(int32_t) * n);
    for (int i = 0; i < n; i++) bara[i] = orc_decode_message(x[i], 2 * N);       /* bootstrap.jl:74 */
    int32_t barb = orc_decode_message(x[n], 2 * N);                              /* bootstrap.jl:75 */
    int32_t* acc = (int32_t*)calloc((size_t)(k + 1) * N, sizeof(int32_t));
    int32_t* tv = (int32_t*)malloc(sizeof(int32_t) * N);
    for (int m = 0; m < N; m++) tv[m] = mu;                                      /* bootstrap.jl:78 */
    orc_mul_by_monomial(tv, -(int64_t)barb, acc + (size_t)k * N, N);             /* bootstrap.jl:54-56 */
    orc_blind_rotate(C, acc, bara, route, n);
    orc_tlwe_extract(acc, k, N, out);
    free(bara); free(acc); free(tv);
}
/* keyswitch.jl:45-80 */
void orc_keyswitch_raw(const int32_t* ksk, int Nk, int n, int t, int basebit, const int32_t* in, int32_t* out) {
    int base = 1 << basebit;
    uint32_t mask = (uint32_t)base - 1;
    uint32_t prec_offset = 1u << (32 - (1 + basebit * t));                       /* keyswitch.jl:58 */
    for (int c = 0; c < n; c++) out[c] = 0;
    out[n] = in[Nk];                                                             /* keyswitch.jl:50 */
    for (int i = 0; i < Nk; i++) {
        int32_t aibar = (int32_t)((uint32_t)in[i] + prec_offset);
        for (int j = 0; j < t; j++) {
            uint32_t d = ((uint32_t)(aibar >> (32 - (j + 1) * basebit))) & mask;
            if (!d) continue;
            const int32_t* row = ksk + (((size_t)i * t + j) * (base - 1) + (d - 1)) * (n + 1);
            for (int c = 0; c <= n; c++) out[c] = (int32_t)((uint32_t)out[c] - (uint32_t)row[c]);
        }
    }
}
void orc_keyswitch(const orc_ctx* C, const int32_t* in, int32_t* out) {
    orc_keyswitch_raw(C->ksk, C->P.N * C->P.k, C->P.n, C->P.t, C->P.basebit, in, out);
}
/* bootstrap.jl:92-95 */
void orc_bootstrap(const orc_ctx* C, int32_t mu, const int32_t* x, int32_t* out, int route) {
    int Nk = C->P.N * C->P.k;
    int32_t* u = (int32_t*)malloc(sizeof(int32_t) * (Nk + 1));
    orc_bootstrap_wo_ks(C, mu, x, u, route);
    orc_keyswitch(C, u, out);
    free(u);
}

/* ------------------------------------------------------------------ L4: gates (gates.jl) */
/* prologue = (0, cb) + ka*x + kb*y; returns 0 for ops without that shape */
static int gate_coeffs(int op, int32_t* cb, int32_t* ka, int32_t* kb) {
    int32_t e8 = orc_encode_message(1, 8), m8 = orc_encode_message(-1, 8);
    int32_t e4 = orc_encode_message(1, 4), m4 = orc_encode_message(-1, 4);
    switch (op) {
        case ORC_NAND:  *cb = e8; *ka = -1; *kb = -1; return 1;   /* gates.jl:16  */
        case ORC_OR:    *cb = e8; *ka = 1;  *kb = 1;  return 1;   /* gates.jl:28  */
        case ORC_AND:   *cb = m8; *ka = 1;  *kb = 1;  return 1;   /* gates.jl:40  */
        case ORC_XOR:   *cb = e4; *ka = 2;  *kb = 2;  return 1;   /* gates.jl:52  */
        case ORC_XNOR:  *cb = m4; *ka = -2; *kb = -2; return 1;   /* gates.jl:64  */
        case ORC_NOR:   *cb = m8; *ka = -1; *kb = -1; return 1;   /* gates.jl:103 */
        case ORC_ANDNY: *cb = m8; *ka = -1; *kb = 1;  return 1;   /* gates.jl:115 */
        case ORC_ANDYN: *cb = m8; *ka = 1;  *kb = -1; return 1;   /* gates.jl:127 */
        case ORC_ORNY:  *cb = e8; *ka = -1; *kb = 1;  return 1;   /* gates.jl:139 */
        case ORC_ORYN:  *cb = e8; *ka = 1;  *kb = -1; return 1;   /* gates.jl:151 */
        default: return 0;
    }
}
static void lin_comb(int32_t cb, int32_t ka, const int32_t* x, int32_t kb, const int32_t* y, int n, int32_t* out) {
    for (int c = 0; c <= n; c++)
        out[c] = (int32_t)((uint32_t)ka * (uint32_t)x[c] + (uint32_t)kb * (uint32_t)y[c]);
    out[n] = (int32_t)((uint32_t)out[n] + (uint32_t)cb);
}
void orc_gate_prologue(int op, const int32_t* x, const int32_t* y, int n, int32_t* out) {
    int32_t cb = 0, ka = 0, kb = 0;
    if (gate_coeffs(op, &cb, &ka, &kb)) lin_comb(cb, ka, x, kb, y, n, out);
    else memset(out, 0, sizeof(int32_t) * (n + 1));
}
static void gate_one(const orc_ctx* C, int op, const int32_t* x, const int32_t* y, const int32_t* z, int32_t* out, int route) {
    int n = C->P.n, Nk = C->P.N * C->P.k;
    int32_t mu = orc_encode_message(1, 8);
    int32_t cb, ka, kb;
    if (gate_coeffs(op, &cb, &ka, &kb)) {
        int32_t* lin = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        lin_comb(cb, ka, x, kb, y, n, lin);
        orc_bootstrap(C, mu, lin, out, route);
        free(lin);
    } else if (op == ORC_NOT) {                                                  /* gates.jl:76-79 */
        for (int c = 0; c <= n; c++) out[c] = (int32_t)(0u - (uint32_t)x[c]);
    } else if (op == ORC_CONSTANT) {                                             /* gates.jl:91-93 */
        for (int c = 0; c < n; c++) out[c] = 0;
        out[n] = orc_encode_message(x && x[0] ? 1 : -1, 8);
    } else if (op == ORC_MUX) {                                                  /* gates.jl:163-177 */
        int32_t* lin = (int32_t*)malloc(sizeof(int32_t) * (n + 1));
        int32_t* u1 = (int32_t*)malloc(sizeof(int32_t) * (Nk + 1) * 2);
        int32_t* u2 = u1 + (Nk + 1);
        lin_comb(orc_encode_message(-1, 8), 1, x, 1, y, n, lin);                 /* :166 */
        orc_bootstrap_wo_ks(C, mu, lin, u1, route);                              /* :167 */
        lin_comb(orc_encode_message(-1, 8), -1, x, 1, z, n, lin);                /* :170 */
        orc_bootstrap_wo_ks(C, mu, lin, u2, route);                              /* :171 */
        for (int c = 0; c <= Nk; c++) u1[c] = (int32_t)((uint32_t)u1[c] + (uint32_t)u2[c]);
        u1[Nk] = (int32_t)((uint32_t)u1[Nk] + (uint32_t)orc_encode_message(1, 8)); /* :174 */
        orc_keyswitch(C, u1, out);                                               /* :176 */
        free(lin); free(u1);
    }
}
void orc_gate_batch(const orc_ctx* C, int op, const int32_t* x, const int32_t* y, const int32_t* z,
                    int32_t* out, size_t count, int route, int nthreads) {
    size_t w = (size_t)C->P.n + 1;
    if (nthreads < 1) nthreads = 1;
    #pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (size_t g = 0; g < count; g++)
        gate_one(C, op, x ? x + g * w : NULL, y ? y + g * w : NULL, z ? z + g * w : NULL, out + g * w, route);
}

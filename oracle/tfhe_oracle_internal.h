/* tfhe_oracle_internal.h — shared between the single-key and multi-key oracle sources.
 * CPU ORACLE: test infrastructure, not product code (see tfhe_oracle.h). */
#ifndef TFHE_ORACLE_INTERNAL_H
#define TFHE_ORACLE_INTERNAL_H
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "tfhe_oracle.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

typedef double complex cplx;

struct orc_ctx {
    orc_params P;
    const int32_t* bk;   /* borrowed, coefficient domain */
    const int32_t* ksk;  /* borrowed */
    cplx* bk_fft;        /* owned: forward_transform.(bk), bootstrap.jl:12 */
};

int32_t orc_rng_torus(orc_rng* r);
int32_t orc_rng_bit(orc_rng* r);
void orc_warm_plan(int N);
void orc_forward_cplx(const int32_t* c, cplx* out, int N);
void orc_inverse_cplx(cplx* in, int32_t* out, int N);
void orc_tlwe_encrypt_zero(orc_rng* rng, double alpha, const int32_t* tlwe_key, int k, int N, int32_t* out);
void orc_keyswitch_raw(const int32_t* ksk, int Nk, int n, int t, int basebit, const int32_t* in, int32_t* out);

#endif

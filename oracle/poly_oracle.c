/* poly_oracle.c -- CPU statement of the caller-side steps around the c64 transform.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * What the reference pins and what it does not.  The reference crate stops at the transform: "the only operations
 * that are performed in the Fourier domain are elementwise" (README.md:10-17), and the integer <-> floating-point
 * conversion, the fold of N real coefficients into N/2 complex points and the negacyclic twist live in its caller
 * (TFHE-rs, not under /root/reference).  What the reference itself shows of these steps:
 *   - the fold: re = coeff[i], im = coeff[i + N/2]                          src/fft128/mod.rs:2006-2016
 *   - the product between fwd and inv, element-wise, then a scaling         src/fft128/mod.rs:2033-2047
 *   - c64 arithmetic of the caller is num_complex's (no FMA)                src/lib.rs:84
 *   - sin/cos come from sincospi64                                          src/fft_simd.rs:237-296
 * The functions below fix the remaining choices for this library (twist table from sincospi64, product without FMA,
 * round-half-away like Rust's f64::round, torus scaling by 2^-64 / 2^64) and are the statement the CUDA kernels are
 * compared with bit for bit; end-to-end correctness is pinned independently by the exact integer schoolbook product
 * in the tests.  "Parity unpinned by the reference" for these steps -- pinned by exact integer arithmetic instead.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* twist[j] = e^{+i pi j / (2 n)} (cos, sin via sincospi64); untwist[j] = conj(twist[j]) / n (exact: n = 2^k) */
void orc_poly_twist_tables(size_t n, oc64 *twist, oc64 *untwist)
{
    const double inv_n = 1.0 / (double)n;
    for (size_t j = 0; j < n; j++) {
        double s, c;
        orc_sincospi64((double)j / (double)(2 * n), &s, &c);
        twist[j].re = c;
        twist[j].im = s;
        untwist[j].re = c * inv_n;
        untwist[j].im = -s * inv_n;
    }
}

static inline oc64 mul_nc(oc64 x, oc64 y) /* num_complex `*`: four products, one subtraction, one addition */
{
    oc64 r;
    r.re = x.re * y.re - x.im * y.im;
    r.im = x.re * y.im + x.im * y.re;
    return r;
}

/* poly: 2n signed 64-bit coefficients (torus mode: the u64 torus element reinterpreted as i64, scaled by 2^-64).
 * out[j] = (f64(poly[j]) * scale + i f64(poly[j + n]) * scale) * twist[j] */
void orc_poly_fold_twist(size_t n, int torus, const int64_t *poly, const oc64 *twist, oc64 *out)
{
    const double scale = torus ? 0x1p-64 : 1.0;
    for (size_t j = 0; j < n; j++) {
        oc64 z;
        z.re = (double)poly[j] * scale;
        z.im = (double)poly[j + n] * scale;
        out[j] = mul_nc(z, twist[j]);
    }
}

static inline uint64_t to_integer(double x) /* round half away from zero (f64::round), then a saturating cast */
{
    const double r = round(x);
    if (r >= 0x1p63) return (uint64_t)INT64_MAX;
    if (r < -0x1p63) return (uint64_t)INT64_MIN;
    if (r != r) return 0;
    return (uint64_t)(int64_t)r;
}
static inline uint64_t to_torus(double x) /* fractional part in [-1/2, 1/2] scaled to 2^64, modulo 2^64 */
{
    double f = x - round(x); /* 0 for every |x| >= 2^52 */
    f = round(f * 0x1p64);
    if (f != f) return 0; /* NaN / infinite input */
    if (f >= 0x1p63) return (uint64_t)1 << 63; /* exactly 2^63: the same torus element as -2^63 */
    return (uint64_t)(int64_t)f;
}

/* z: n Fourier^-1 points (output of the unnormalised inverse).  t = z[j] * untwist[j];
 * out[j] (+)= conv(t.re), out[j + n] (+)= conv(t.im); `accumulate` adds modulo 2^64. */
void orc_poly_untwist_round(size_t n, int torus, int accumulate, const oc64 *z, const oc64 *untwist, int64_t *out)
{
    for (size_t j = 0; j < n; j++) {
        const oc64 t = mul_nc(z[j], untwist[j]);
        const uint64_t re = torus ? to_torus(t.re) : to_integer(t.re);
        const uint64_t im = torus ? to_torus(t.im) : to_integer(t.im);
        if (accumulate) {
            out[j] = (int64_t)((uint64_t)out[j] + re);
            out[j + n] = (int64_t)((uint64_t)out[j + n] + im);
        } else {
            out[j] = (int64_t)re;
            out[j + n] = (int64_t)im;
        }
    }
}

/* out[r] (+)= round(untwist(inv(sum_k fwd(twist(fold(a[r][k]))) (.) b[r or shared][k]))) for rows [lo, hi): the whole
 * negacyclic product step as the composition of the pieces above and the transform oracle. */
struct poly_mul_ctx {
    const orc_unordered_plan *plan;
    size_t n, k;
    int torus, accumulate;
    const int64_t *a;
    const oc64 *b;
    size_t b_row_stride;
    int64_t *out;
    const oc64 *twist, *untwist;
    size_t base_n;
};

static void poly_mul_rows(void *vctx, size_t lo, size_t hi)
{
    const struct poly_mul_ctx *c = (const struct poly_mul_ctx *)vctx;
    const size_t n = c->n;
    oc64 *term = (oc64 *)aligned_alloc(64, n * sizeof(oc64));
    oc64 *acc = (oc64 *)aligned_alloc(64, n * sizeof(oc64));
    oc64 *scratch = (oc64 *)aligned_alloc(64, c->base_n * sizeof(oc64));
    for (size_t r = lo; r < hi; r++) {
        for (size_t k = 0; k < c->k; k++) {
            orc_poly_fold_twist(n, c->torus, c->a + (r * c->k + k) * 2 * n, c->twist, term);
            orc_unordered_fwd(c->plan, term, scratch);
            const oc64 *bk = c->b + r * c->b_row_stride + k * n;
            if (k == 0) {
                orc_c64_pointwise(NULL, (double *)term, (const double *)bk, n);
                memcpy(acc, term, n * sizeof(oc64));
            } else {
                orc_c64_pointwise((double *)acc, (double *)term, (const double *)bk, n);
            }
        }
        orc_unordered_inv(c->plan, acc, scratch);
        orc_poly_untwist_round(n, c->torus, c->accumulate, acc, c->untwist, c->out + r * 2 * n);
    }
    free(term);
    free(acc);
    free(scratch);
}

void orc_poly_mul_batch(const orc_unordered_plan *plan, size_t n, size_t base_n, const int64_t *a, size_t k_terms, const oc64 *b,
                        size_t b_row_stride, int64_t *out, size_t batch, int torus, int accumulate, int threads)
{
    oc64 *twist = (oc64 *)malloc(n * sizeof(oc64)), *untwist = (oc64 *)malloc(n * sizeof(oc64));
    orc_poly_twist_tables(n, twist, untwist);
    struct poly_mul_ctx c = {plan, n, k_terms, torus, accumulate, a, b, b_row_stride, out, twist, untwist, base_n};
    orc_parallel_rows(threads, batch, poly_mul_rows, &c);
    free(twist);
    free(untwist);
}

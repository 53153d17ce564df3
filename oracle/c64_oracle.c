/*
 * c64_oracle.c -- bit-faithful CPU restatement of concrete-fft's c64 transforms.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Compile with
 *   gcc -O2 -ffp-contract=off -mfma
 * so that every fma() below is one fused operation and nothing else is fused.
 *
 * The reference's SIMD paths (AVX2 / AVX-512) compute lane-wise the same IEEE
 * operations as its scalar path (src/x86.rs:51-58 == src/fft_simd.rs:220-233),
 * so this scalar restatement is the reference result on every platform.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* complex helpers: src/fft_simd.rs:106-160 (derived ops), :201-233    */
/* ------------------------------------------------------------------ */

static const double INV_SQRT2 = 0.7071067811865476; /* core::f64::consts::FRAC_1_SQRT_2 */
static const double H1X = 0.9238795325112867;       /* src/fft_simd.rs:50 */
static const double H1Y = -0.38268343236508984;     /* src/fft_simd.rs:52 */

static inline oc64 cx(double re, double im) { oc64 z = { re, im }; return z; }
static inline oc64 cadd_s(oc64 a, oc64 b) { return cx(a.re + b.re, a.im + b.im); }
static inline oc64 csub_s(oc64 a, oc64 b) { return cx(a.re - b.re, a.im - b.im); }

/* src/fft_simd.rs:220-233: w * z, re = fma(a, x, -(b*y)), im = fma(a, y, b*x) */
static inline oc64 cmul_s(oc64 w, oc64 z)
{
    double a = w.re, b = w.im, x = z.re, y = z.im;
    return cx(fma(a, x, -b * y), fma(a, y, b * x));
}

/* src/fft_simd.rs:113-120 */
static inline oc64 mulj_s(int fwd, oc64 z)
{
    return fwd ? cx(-z.im, z.re) : cx(z.im, -z.re);
}

/* src/fft_simd.rs:122-131 */
static inline oc64 mul_e8_s(int fwd, oc64 z)
{
    oc64 t = cadd_s(z, mulj_s(fwd, z));
    return cx(INV_SQRT2 * t.re, INV_SQRT2 * t.im);
}
static inline oc64 mul_ne8_s(int fwd, oc64 z) { return mul_e8_s(!fwd, z); }

/* src/fft_simd.rs:133-159 */
static inline oc64 mul_e16_s(int fwd, oc64 z) { return cmul_s(cx(H1X, fwd ? H1Y : -H1Y), z); }
static inline oc64 mul_e17_s(int fwd, oc64 z) { return cmul_s(cx(-H1Y, fwd ? -H1X : H1X), z); }
static inline oc64 mul_ne16_s(int fwd, oc64 z) { return mul_e16_s(!fwd, z); }
static inline oc64 mul_ne17_s(int fwd, oc64 z) { return mul_e17_s(!fwd, z); }

#define CT oc64
#define X(f) f##_s
#include "butterflies.inc"
#undef CT
#undef X
/* ---- AVX2 + FMA instantiation: a register holds two interleaved complex numbers, the reference's
 * c64x2 (src/x86.rs:4-74: add/sub lane-wise, mul = fmaddsub(aa, xy, bb * yx), swap = permute 0b0101).
 * Used only by the -O3 -march=x86-64-v3 build (the timed CPU baseline); same bits as the scalar code. */
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_HAVE_AVX2 1
typedef __m256d cv;
static inline cv cadd_v(cv a, cv b) { return _mm256_add_pd(a, b); }
static inline cv csub_v(cv a, cv b) { return _mm256_sub_pd(a, b); }
static inline cv cmul_v(cv w, cv z)
{
    cv yx = _mm256_permute_pd(z, 0x5);
    cv aa = _mm256_unpacklo_pd(w, w), bb = _mm256_unpackhi_pd(w, w);
    return _mm256_fmaddsub_pd(aa, z, _mm256_mul_pd(bb, yx));
}
static inline cv mulj_v(int fwd, cv z)
{
    cv sw = _mm256_permute_pd(z, 0x5); /* (im, re) */
    const cv neg_even = _mm256_set_pd(0.0, -0.0, 0.0, -0.0), neg_odd = _mm256_set_pd(-0.0, 0.0, -0.0, 0.0);
    return _mm256_xor_pd(sw, fwd ? neg_even : neg_odd); /* fwd: (-im, re); else (im, -re) */
}
static inline cv mul_e8_v(int fwd, cv z)
{
    return _mm256_mul_pd(_mm256_set1_pd(INV_SQRT2), cadd_v(z, mulj_v(fwd, z)));
}
static inline cv mul_ne8_v(int fwd, cv z) { return mul_e8_v(!fwd, z); }
static inline cv cconst_v(double re, double im) { return _mm256_set_pd(im, re, im, re); }
static inline cv mul_e16_v(int fwd, cv z) { return cmul_v(cconst_v(H1X, fwd ? H1Y : -H1Y), z); }
static inline cv mul_e17_v(int fwd, cv z) { return cmul_v(cconst_v(-H1Y, fwd ? -H1X : H1X), z); }
static inline cv mul_ne16_v(int fwd, cv z) { return mul_e16_v(!fwd, z); }
static inline cv mul_ne17_v(int fwd, cv z) { return mul_e17_v(!fwd, z); }
#define CT cv
#define X(f) f##_v
#include "butterflies.inc"
#undef CT
#undef X
static inline cv ld2(const oc64 *p) { return _mm256_loadu_pd((const double *)p); }
static inline cv ld11(const oc64 *p0, const oc64 *p1)
{
    return _mm256_set_m128d(_mm_loadu_pd((const double *)p1), _mm_loadu_pd((const double *)p0));
}
static inline cv splat1(const oc64 *p) { return _mm256_broadcast_pd((const __m128d *)p); }
static inline void st2(oc64 *p, cv v) { _mm256_storeu_pd((double *)p, v); }
static inline void st11(oc64 *p0, oc64 *p1, cv v)
{
    _mm_storeu_pd((double *)p0, _mm256_castpd256_pd128(v));
    _mm_storeu_pd((double *)p1, _mm256_extractf128_pd(v, 1));
}
#endif

/* the scalar drivers below use the unsuffixed names */
#define cadd cadd_s
#define csub csub_s
#define cmul cmul_s
#define bfR bfR_s

/* ------------------------------------------------------------------ */
/* sincospi64 and twiddle tables                                       */
/* ------------------------------------------------------------------ */

/* src/fft_simd.rs:237-296 */
void orc_sincospi64(double a, double *s_out, double *c_out)
{
    double az = a * 0.0;
    a = (fabs(a) < 9007199254740992.0) ? a : az;

    double r = round(a + a); /* Rust f64::round: half away from zero == C round() */
    int64_t i = (int64_t)r;
    double t = fma(-0.5, r, a);
    double s = t * t;

    r = -1.0369917389758117e-4;
    r = fma(r, s, 1.9294935641298806e-3);
    r = fma(r, s, -2.5806887942825395e-2);
    r = fma(r, s, 2.3533063028328211e-1);
    r = fma(r, s, -1.3352627688538006e+0);
    r = fma(r, s, 4.0587121264167623e+0);
    r = fma(r, s, -4.9348022005446790e+0);
    double c = fma(r, s, 1.0000000000000000e+0);

    r = 4.6151442520157035e-4;
    r = fma(r, s, -7.3700183130883555e-3);
    r = fma(r, s, 8.2145868949323936e-2);
    r = fma(r, s, -5.9926452893214921e-1);
    r = fma(r, s, 2.5501640398732688e+0);
    r = fma(r, s, -5.1677127800499516e+0);
    s = s * t;
    r = r * s;
    s = fma(t, 3.1415926535897931e+0, r);

    if (i & 2) {
        s = 0.0 - s;
        c = 0.0 - c;
    }
    if (i & 1) {
        double tt = 0.0 - s;
        s = c;
        c = tt;
    }
    if (a == floor(a))
        s = az;
    *s_out = s;
    *c_out = c;
}

/* src/fft_simd.rs:298-321 */
void orc_init_wt(size_t r, size_t n, oc64 *w, oc64 *w_inv)
{
    if (n < r)
        return;
    size_t nr = n / r;
    double theta = -2.0 / (double)n;
    for (size_t i = 0; i < 2 * n; i++)
        w[i] = cx(NAN, NAN);
    for (size_t p = 0; p < nr; p++) {
        for (size_t k = 1; k < r; k++) {
            double s, c;
            orc_sincospi64(theta * (double)(k * p), &s, &c);
            w[p + k * nr] = cx(c, s);
            w[n + r * p + k] = cx(c, s);
            w_inv[p + k * nr] = cx(c, -s);
            w_inv[n + r * p + k] = cx(c, -s);
        }
    }
}

/* ------------------------------------------------------------------ */
/* ordered (Stockham autosort) stages                                  */
/* ------------------------------------------------------------------ */

/* DIF core, radix R, stride s: e.g. src/dif4.rs:118-168, src/dif16.rs:449-623.
 * reads x[q + s(p + m k)], writes y[q + s(R p + k)] = w[R p s + k] * DFT_R(x)_k */
static inline __attribute__((always_inline)) void dif_core_impl(const int R, const int fwd, size_t n, size_t s, const oc64 *x, oc64 *y, const oc64 *w)
{
    size_t m = n / ((size_t)R * s);
    for (size_t p = 0; p < m; p++) {
        const oc64 *wp = w + (size_t)R * p * s;
        for (size_t q = 0; q < s; q++) {
            oc64 v[16];
            for (int k = 0; k < R; k++)
                v[k] = x[q + s * (p + m * (size_t)k)];
            bfR(R, fwd, v);
            y[q + s * ((size_t)R * p)] = v[0];
            for (int k = 1; k < R; k++)
                y[q + s * ((size_t)R * p + (size_t)k)] = cmul(wp[k], v[k]);
        }
    }
}

#ifdef ORC_HAVE_AVX2
/* two butterflies per iteration: pairs of q (s >= 2, twiddle splat) or pairs of p (s == 1) */
static inline __attribute__((always_inline)) void dif_core_impl_v(const int R, const int fwd, size_t n, size_t s, const oc64 *x, oc64 *y, const oc64 *w)
{
    size_t m = n / ((size_t)R * s);
    cv v[16];
    if (s >= 2) {
        for (size_t p = 0; p < m; p++) {
            const oc64 *wp = w + (size_t)R * p * s;
            for (size_t q = 0; q < s; q += 2) {
                for (int k = 0; k < R; k++) v[k] = ld2(&x[q + s * (p + m * (size_t)k)]);
                bfR_v(R, fwd, v);
                st2(&y[q + s * ((size_t)R * p)], v[0]);
                for (int k = 1; k < R; k++) st2(&y[q + s * ((size_t)R * p + (size_t)k)], cmul_v(splat1(&wp[k]), v[k]));
            }
        }
    } else {
        for (size_t p = 0; p < m; p += 2) {
            const oc64 *w0 = w + (size_t)R * p, *w1 = w0 + R;
            oc64 *y0 = y + (size_t)R * p, *y1 = y0 + R;
            for (int k = 0; k < R; k++) v[k] = ld2(&x[p + m * (size_t)k]);
            bfR_v(R, fwd, v);
            st11(&y0[0], &y1[0], v[0]);
            for (int k = 1; k < R; k++) st11(&y0[k], &y1[k], cmul_v(ld11(&w0[k], &w1[k]), v[k]));
        }
    }
}
#define dif_core_impl(R, F, n, s, x, y, w) (((n) / ((size_t)(R) * (s))) >= 2 || (s) >= 2 ? dif_core_impl_v(R, F, n, s, x, y, w) : dif_core_impl(R, F, n, s, x, y, w))
#endif

/* radix / direction become compile-time constants in each arm (speed only; same arithmetic) */
static void dif_core(int R, int fwd, size_t n, size_t s, const oc64 *x, oc64 *y, const oc64 *w)
{
#define ARM(r) case r: if (fwd) dif_core_impl(r, 1, n, s, x, y, w); else dif_core_impl(r, 0, n, s, x, y, w); break
    switch (R) { ARM(2); ARM(4); ARM(8); default: if (fwd) dif_core_impl(16, 1, n, s, x, y, w); else dif_core_impl(16, 0, n, s, x, y, w); break; }
#undef ARM
}

/* DIT core, radix R, stride s: e.g. src/dit4.rs:96-145, src/dit16.rs:366-548.
 * reads y[q + s(R p + k)] * w[R p s + k], writes x[q + s(p + m k)] = DFT_R(.)_k */
static inline __attribute__((always_inline)) void dit_core_impl(const int R, const int fwd, size_t n, size_t s, oc64 *x, const oc64 *y, const oc64 *w)
{
    size_t m = n / ((size_t)R * s);
    for (size_t p = 0; p < m; p++) {
        const oc64 *wp = w + (size_t)R * p * s;
        for (size_t q = 0; q < s; q++) {
            oc64 v[16];
            v[0] = y[q + s * ((size_t)R * p)];
            for (int k = 1; k < R; k++)
                v[k] = cmul(wp[k], y[q + s * ((size_t)R * p + (size_t)k)]);
            bfR(R, fwd, v);
            for (int k = 0; k < R; k++)
                x[q + s * (p + m * (size_t)k)] = v[k];
        }
    }
}

/* radix / direction become compile-time constants in each arm (speed only; same arithmetic) */
#ifdef ORC_HAVE_AVX2
static inline __attribute__((always_inline)) void dit_core_impl_v(const int R, const int fwd, size_t n, size_t s, oc64 *x, const oc64 *y, const oc64 *w)
{
    size_t m = n / ((size_t)R * s);
    cv v[16];
    if (s >= 2) {
        for (size_t p = 0; p < m; p++) {
            const oc64 *wp = w + (size_t)R * p * s;
            for (size_t q = 0; q < s; q += 2) {
                v[0] = ld2(&y[q + s * ((size_t)R * p)]);
                for (int k = 1; k < R; k++) v[k] = cmul_v(splat1(&wp[k]), ld2(&y[q + s * ((size_t)R * p + (size_t)k)]));
                bfR_v(R, fwd, v);
                for (int k = 0; k < R; k++) st2(&x[q + s * (p + m * (size_t)k)], v[k]);
            }
        }
    } else {
        for (size_t p = 0; p < m; p += 2) {
            const oc64 *w0 = w + (size_t)R * p, *w1 = w0 + R;
            const oc64 *y0 = y + (size_t)R * p, *y1 = y0 + R;
            v[0] = ld11(&y0[0], &y1[0]);
            for (int k = 1; k < R; k++) v[k] = cmul_v(ld11(&w0[k], &w1[k]), ld11(&y0[k], &y1[k]));
            bfR_v(R, fwd, v);
            for (int k = 0; k < R; k++) st2(&x[p + m * (size_t)k], v[k]);
        }
    }
}
#define dit_core_impl(R, F, n, s, x, y, w) (((n) / ((size_t)(R) * (s))) >= 2 || (s) >= 2 ? dit_core_impl_v(R, F, n, s, x, y, w) : dit_core_impl(R, F, n, s, x, y, w))
#endif

static void dit_core(int R, int fwd, size_t n, size_t s, oc64 *x, const oc64 *y, const oc64 *w)
{
#define ARM(r) case r: if (fwd) dit_core_impl(r, 1, n, s, x, y, w); else dit_core_impl(r, 0, n, s, x, y, w); break
    switch (R) { ARM(2); ARM(4); ARM(8); default: if (fwd) dit_core_impl(16, 1, n, s, x, y, w); else dit_core_impl(16, 0, n, s, x, y, w); break; }
#undef ARM
}

/* terminal twiddle-free pass: e.g. src/dif4.rs:217-244; dst may alias src */
static inline __attribute__((always_inline)) void end_stage_impl(const int R, const int fwd, size_t n, const oc64 *src, oc64 *dst)
{
    size_t part = n / (size_t)R;
    for (size_t j = 0; j < part; j++) {
        oc64 v[16];
        for (int k = 0; k < R; k++)
            v[k] = src[(size_t)k * part + j];
        bfR(R, fwd, v);
        for (int k = 0; k < R; k++)
            dst[(size_t)k * part + j] = v[k];
    }
}

/* radix / direction become compile-time constants in each arm (speed only; same arithmetic) */
#ifdef ORC_HAVE_AVX2
static inline __attribute__((always_inline)) void end_stage_impl_v(const int R, const int fwd, size_t n, const oc64 *src, oc64 *dst)
{
    size_t part = n / (size_t)R;
    cv v[16];
    for (size_t j = 0; j < part; j += 2) {
        for (int k = 0; k < R; k++) v[k] = ld2(&src[(size_t)k * part + j]);
        bfR_v(R, fwd, v);
        for (int k = 0; k < R; k++) st2(&dst[(size_t)k * part + j], v[k]);
    }
}
#define end_stage_impl(R, F, n, src, dst) ((n) / (size_t)(R) >= 2 ? end_stage_impl_v(R, F, n, src, dst) : end_stage_impl(R, F, n, src, dst))
#endif

static void end_stage(int R, int fwd, size_t n, const oc64 *src, oc64 *dst)
{
#define ARM(r) case r: if (fwd) end_stage_impl(r, 1, n, src, dst); else end_stage_impl(r, 0, n, src, dst); break
    switch (R) { ARM(2); ARM(4); ARM(8); default: if (fwd) end_stage_impl(16, 1, n, src, dst); else end_stage_impl(16, 0, n, src, dst); break; }
#undef ARM
}

static int algo_radix(int algo)
{
    switch (algo) {
    case ORC_DIF2: case ORC_DIT2: return 2;
    case ORC_DIF4: case ORC_DIT4: return 4;
    case ORC_DIF8: case ORC_DIT8: return 8;
    default: return 16;
    }
}
static int algo_is_dit(int algo) { return algo & 1; }
static unsigned ilog2(size_t n) { unsigned b = 0; while ((n >> b) > 1) b++; return b; }

/* type-level recursion of e.g. src/dif4.rs:246-303 unrolled into a loop */
static void dif_run(int R, int fwd, size_t n, oc64 *buf, oc64 *scratch, const oc64 *w)
{
    unsigned rho = ilog2((size_t)R), bits = ilog2(n);
    oc64 *x = buf, *y = scratch;
    int write_to_x = 1;
    size_t s = 1;
    while (bits > rho) {
        dif_core(R, fwd, n, s, x, y, w);
        oc64 *t = x; x = y; y = t;
        write_to_x = !write_to_x;
        s *= (size_t)R;
        bits -= rho;
    }
    end_stage(1 << bits, fwd, n, x, write_to_x ? x : y);
}

/* e.g. src/dit4.rs:223-280: recursion first (terminal pass), cores on the way back */
static void dit_rec(int R, int fwd, size_t n, unsigned bits, int read_from_x, size_t s,
                    oc64 *x, oc64 *y, const oc64 *w)
{
    unsigned rho = ilog2((size_t)R);
    if (bits <= rho) {
        end_stage(1 << bits, fwd, n, read_from_x ? x : y, x);
        return;
    }
    dit_rec(R, fwd, n, bits - rho, !read_from_x, s * (size_t)R, y, x, w);
    dit_core(R, fwd, n, s, x, y, w);
}

/* fn-pointer contract of src/lib.rs:160-226; w_tab has 2n entries, second half used by
 * the scalar path.  DIF2/DIT2 use fwd = true for both directions (src/dif2.rs:188-204). */
static void ordered_run(int algo, int fwd, size_t n, oc64 *buf, oc64 *scratch, const oc64 *w_tab)
{
    if (n == 1)
        return; /* src/ordered.rs:210-212 */
    int R = algo_radix(algo);
    const oc64 *w = w_tab + n;
    if (algo_is_dit(algo))
        dit_rec(R, fwd, n, ilog2(n), 1, 1, buf, scratch, w);
    else
        dif_run(R, fwd, n, buf, scratch, w);
}

struct orc_ordered_plan {
    size_t n;
    int algo;
    oc64 *tw, *tw_inv; /* 2n each */
};

static int is_pow2(size_t n) { return n && !(n & (n - 1)); }

orc_ordered_plan *orc_ordered_plan_new(size_t n, int algo)
{
    if (!is_pow2(n) || ilog2(n) >= 11 || algo < 0 || algo > 7)
        return NULL; /* src/ordered.rs:243-244 */
    orc_ordered_plan *p = calloc(1, sizeof *p);
    p->n = n;
    p->algo = algo;
    p->tw = calloc(2 * n, sizeof(oc64));
    p->tw_inv = calloc(2 * n, sizeof(oc64));
    orc_init_wt((size_t)algo_radix(algo), n, p->tw, p->tw_inv);
    return p;
}
void orc_ordered_plan_free(orc_ordered_plan *p)
{
    if (!p) return;
    free(p->tw); free(p->tw_inv); free(p);
}
void orc_ordered_fwd(const orc_ordered_plan *p, oc64 *buf, oc64 *scratch)
{
    ordered_run(p->algo, 1, p->n, buf, scratch, p->tw);
}
void orc_ordered_inv(const orc_ordered_plan *p, oc64 *buf, oc64 *scratch)
{
    ordered_run(p->algo, 0, p->n, buf, scratch, p->tw_inv);
}

/* ------------------------------------------------------------------ */
/* unordered plan                                                      */
/* ------------------------------------------------------------------ */

static int top_radix(size_t n, size_t base_n)
{
    /* src/unordered.rs:407-413 */
    return n == 2 * base_n ? 2 : (n == 4 * base_n ? 4 : 8);
}

/* src/unordered.rs:349-389 with complex_per_reg = 1 */
static void init_twiddles(size_t n, size_t base_n, size_t base_r, oc64 *w, size_t w_len,
                          oc64 *w_inv, size_t w_inv_len)
{
    double theta = 2.0 / (double)n;
    if (n <= base_n) {
        (void)w_len; (void)w_inv_len;
        orc_init_wt(base_r, n, w, w_inv);
        return;
    }
    size_t r = (size_t)top_radix(n, base_n);
    size_t m = n / r;
    size_t lvl = (r - 1) * m;
    oc64 *w_next = w + lvl;
    oc64 *w_inv_lvl = w_inv + (w_inv_len - lvl);
    for (size_t p = 0; p < m; p++) {
        for (size_t k = 1; k < r; k++) {
            double sk, ck;
            orc_sincospi64(theta * (double)(k * p), &sk, &ck);
            size_t idx = (r - 1) * p + (k - 1);
            w[idx] = cx(ck, -sk);
            w_inv_lvl[idx] = cx(ck, sk);
        }
    }
    init_twiddles(n / r, base_n, base_r, w_next, w_len - lvl, w_inv, w_inv_len - lvl);
}

static size_t brev(unsigned nbits, size_t i)
{
    size_t r = 0;
    for (unsigned b = 0; b < nbits; b++)
        r |= ((i >> b) & 1) << (nbits - 1 - b);
    return r;
}

/* src/unordered.rs:222-293 (fwd_process_x{2,4,8}) */
static inline __attribute__((always_inline)) void fwd_top_stage_impl(const int r, size_t n, oc64 *z, const oc64 *w)
{
    size_t m = n / (size_t)r;
    unsigned rb = ilog2((size_t)r);
    for (size_t p = 0; p < m; p++) {
        oc64 v[8];
        for (int k = 0; k < r; k++)
            v[k] = z[p + m * (size_t)k];
        bfR(r, 1, v);
        const oc64 *wp = w + (size_t)(r - 1) * p;
        z[p] = v[0];
        for (int k = 1; k < r; k++)
            z[p + m * brev(rb, (size_t)k)] = cmul(wp[k - 1], v[k]);
    }
}

#ifdef ORC_HAVE_AVX2
/* pairs of p: n / r >= base_n >= 32 here, so it is always even */
static inline __attribute__((always_inline)) void fwd_top_stage_impl_v(const int r, size_t n, oc64 *z, const oc64 *w)
{
    size_t m = n / (size_t)r;
    unsigned rb = ilog2((size_t)r);
    cv v[8];
    for (size_t p = 0; p < m; p += 2) {
        const oc64 *w0 = w + (size_t)(r - 1) * p, *w1 = w0 + (r - 1);
        for (int k = 0; k < r; k++) v[k] = ld2(&z[p + m * (size_t)k]);
        bfR_v(r, 1, v);
        st2(&z[p], v[0]);
        for (int k = 1; k < r; k++) st2(&z[p + m * brev(rb, (size_t)k)], cmul_v(ld11(&w0[k - 1], &w1[k - 1]), v[k]));
    }
}
static inline __attribute__((always_inline)) void inv_top_stage_impl_v(const int r, size_t n, oc64 *z, const oc64 *w)
{
    size_t m = n / (size_t)r;
    unsigned rb = ilog2((size_t)r);
    cv v[8];
    for (size_t p = 0; p < m; p += 2) {
        const oc64 *w0 = w + (size_t)(r - 1) * p, *w1 = w0 + (r - 1);
        v[0] = ld2(&z[p]);
        for (int k = 1; k < r; k++) v[k] = cmul_v(ld11(&w0[k - 1], &w1[k - 1]), ld2(&z[p + m * brev(rb, (size_t)k)]));
        bfR_v(r, 0, v);
        for (int k = 0; k < r; k++) st2(&z[p + m * (size_t)k], v[k]);
    }
}
#define fwd_top_stage_impl fwd_top_stage_impl_v
#endif

static void fwd_top_stage(int r, size_t n, oc64 *z, const oc64 *w)
{
    switch (r) {
    case 2: fwd_top_stage_impl(2, n, z, w); break;
    case 4: fwd_top_stage_impl(4, n, z, w); break;
    default: fwd_top_stage_impl(8, n, z, w); break;
    }
}

/* src/unordered.rs:232-293 (inv_process_x{2,4,8}) */
static inline __attribute__((always_inline)) void inv_top_stage_impl(const int r, size_t n, oc64 *z, const oc64 *w)
{
    size_t m = n / (size_t)r;
    unsigned rb = ilog2((size_t)r);
    for (size_t p = 0; p < m; p++) {
        oc64 v[8];
        const oc64 *wp = w + (size_t)(r - 1) * p;
        v[0] = z[p];
        for (int k = 1; k < r; k++)
            v[k] = cmul(wp[k - 1], z[p + m * brev(rb, (size_t)k)]);
        bfR(r, 0, v);
        for (int k = 0; k < r; k++)
            z[p + m * (size_t)k] = v[k];
    }
}

#ifdef ORC_HAVE_AVX2
#define inv_top_stage_impl inv_top_stage_impl_v
#endif
static void inv_top_stage(int r, size_t n, oc64 *z, const oc64 *w)
{
    switch (r) {
    case 2: inv_top_stage_impl(2, n, z, w); break;
    case 4: inv_top_stage_impl(4, n, z, w); break;
    default: inv_top_stage_impl(8, n, z, w); break;
    }
}

struct orc_unordered_plan {
    size_t n, base_n;
    int base_algo;
    oc64 *tw, *tw_inv;   /* n + base_n each */
    oc64 *monomial_tw;   /* n */
    size_t *indices;     /* n */
};

/* src/unordered.rs:391-439 */
static void fwd_depth(const orc_unordered_plan *pl, oc64 *z, size_t n, const oc64 *w, oc64 *scratch)
{
    if (n == pl->base_n) {
        ordered_run(pl->base_algo, 1, n, z, scratch, w);
        return;
    }
    int r = top_radix(n, pl->base_n);
    size_t m = n / (size_t)r;
    fwd_top_stage(r, n, z, w);
    const oc64 *w_tail = w + (size_t)(r - 1) * m;
    for (int c = 0; c < r; c++)
        fwd_depth(pl, z + (size_t)c * m, m, w_tail, scratch);
}

/* src/unordered.rs:441-489; w_len = length of the table slice seen at this level */
static void inv_depth(const orc_unordered_plan *pl, oc64 *z, size_t n, const oc64 *w, size_t w_len,
                      oc64 *scratch)
{
    if (n == pl->base_n) {
        ordered_run(pl->base_algo, 0, n, z, scratch, w);
        return;
    }
    int r = top_radix(n, pl->base_n);
    size_t m = n / (size_t)r;
    size_t head_len = w_len - (size_t)(r - 1) * m;
    for (int c = 0; c < r; c++)
        inv_depth(pl, z + (size_t)c * m, m, w, head_len, scratch);
    inv_top_stage(r, n, z, w + head_len);
}

orc_unordered_plan *orc_unordered_plan_new(size_t n, int base_algo, size_t base_n)
{
    /* src/unordered.rs:659-671 */
    if (!is_pow2(n) || !is_pow2(base_n) || base_n > n || base_algo < 0 || base_algo > 7)
        return NULL;
    if (base_n != n && base_n < 32)
        return NULL;
    if (ilog2(base_n) > 10)
        return NULL;
    orc_unordered_plan *p = calloc(1, sizeof *p);
    p->n = n;
    p->base_n = base_n;
    p->base_algo = base_algo;
    size_t len = n + base_n;
    p->tw = malloc(len * sizeof(oc64));
    p->tw_inv = malloc(len * sizeof(oc64));
    for (size_t i = 0; i < len; i++)
        p->tw[i] = p->tw_inv[i] = cx(NAN, NAN);
    init_twiddles(n, base_n, (size_t)algo_radix(base_algo), p->tw, len, p->tw_inv, len);

    /* src/unordered.rs:714-728 */
    p->monomial_tw = malloc(n * sizeof(oc64));
    double theta = -2.0 / (double)n;
    for (size_t i = 0; i < n; i++) {
        double s, c;
        orc_sincospi64(theta * (double)i, &s, &c);
        p->monomial_tw[i] = cx(c, s);
    }
    p->indices = malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; i++)
        p->indices[i] = orc_bit_rev_twice_inv(ilog2(n), ilog2(base_n), i);
    return p;
}

void orc_unordered_plan_free(orc_unordered_plan *p)
{
    if (!p) return;
    free(p->tw); free(p->tw_inv); free(p->monomial_tw); free(p->indices); free(p);
}

const oc64 *orc_unordered_twiddles(const orc_unordered_plan *p, int inverse)
{
    return inverse ? p->tw_inv : p->tw;
}

void orc_unordered_fwd(const orc_unordered_plan *p, oc64 *buf, oc64 *scratch)
{
    fwd_depth(p, buf, p->n, p->tw, scratch);
}

void orc_unordered_inv(const orc_unordered_plan *p, oc64 *buf, oc64 *scratch)
{
    inv_depth(p, buf, p->n, p->tw_inv, p->n + p->base_n, scratch);
}

/* src/unordered.rs:844-900 */
void orc_unordered_fwd_monomial(const orc_unordered_plan *pl, size_t degree, oc64 *buf)
{
    size_t n = pl->n, mask = n - 1;
    const oc64 *tw = pl->monomial_tw;
    switch (n / pl->base_n) {
    case 1:
        for (size_t i = 0; i < n; i++)
            buf[i] = tw[(i * degree) & mask];
        break;
    case 2:
        for (size_t i = 0; i < n / 2; i++) {
            buf[i] = tw[((2 * i) * degree) & mask];
            buf[n / 2 + i] = tw[((2 * i + 1) * degree) & mask];
        }
        break;
    default:
        for (size_t i = 0; i < n; i++)
            buf[i] = tw[(pl->indices[i] * degree) & mask];
        break;
    }
}

/* batch drivers: rows are independent (one Plan::fwd call per polynomial in the reference) */
struct ubatch { const orc_unordered_plan *p; oc64 *buf; int inverse; };
static void ubatch_rows(void *ctx, size_t lo, size_t hi)
{
    struct ubatch *u = ctx;
    oc64 *scratch = malloc(u->p->base_n * sizeof(oc64));
    for (size_t b = lo; b < hi; b++) {
        if (u->inverse)
            orc_unordered_inv(u->p, u->buf + b * u->p->n, scratch);
        else
            orc_unordered_fwd(u->p, u->buf + b * u->p->n, scratch);
    }
    free(scratch);
}

void orc_unordered_fwd_batch(const orc_unordered_plan *p, oc64 *buf, size_t batch, int threads)
{
    struct ubatch u = { p, buf, 0 };
    orc_parallel_rows(threads, batch, ubatch_rows, &u);
}

void orc_unordered_inv_batch(const orc_unordered_plan *p, oc64 *buf, size_t batch, int threads)
{
    struct ubatch u = { p, buf, 1 };
    orc_parallel_rows(threads, batch, ubatch_rows, &u);
}

/* ------------------------------------------------------------------ */
/* permutation: src/unordered.rs:1039-1059                             */
/* ------------------------------------------------------------------ */

size_t orc_bit_rev(unsigned nbits, size_t i)
{
    return nbits == 0 ? 0 : brev(nbits, i);
}

size_t orc_bit_rev_twice(unsigned nbits, unsigned base_nbits, size_t i)
{
    size_t i_rev = orc_bit_rev(nbits, i);
    size_t bottom_mask = ((size_t)1 << base_nbits) - 1;
    size_t bottom_bits = orc_bit_rev(base_nbits, i_rev & bottom_mask);
    return (i_rev & ~bottom_mask) | bottom_bits;
}

size_t orc_bit_rev_twice_inv(unsigned nbits, unsigned base_nbits, size_t i)
{
    size_t bottom_mask = ((size_t)1 << base_nbits) - 1;
    size_t bottom_bits = orc_bit_rev(base_nbits, i & bottom_mask);
    size_t i_rev = (i & ~bottom_mask) | bottom_bits;
    return orc_bit_rev(nbits, i_rev);
}

/* Fourier-domain element-wise product as a Rust caller writes it on `c64` values (num_complex Mul / Add:
 * each operation individually rounded; this file is built with -ffp-contract=off). */
void orc_c64_pointwise(double *acc, double *a, const double *b, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        const double ar = a[2 * i], ai = a[2 * i + 1], br = b[2 * i], bi = b[2 * i + 1];
        const double re = ar * br - ai * bi, im = ar * bi + ai * br;
        if (acc) {
            acc[2 * i] = acc[2 * i] + re;
            acc[2 * i + 1] = acc[2 * i + 1] + im;
        } else {
            a[2 * i] = re;
            a[2 * i + 1] = im;
        }
    }
}

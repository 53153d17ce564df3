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
static inline oc64 cadd(oc64 a, oc64 b) { return cx(a.re + b.re, a.im + b.im); }
static inline oc64 csub(oc64 a, oc64 b) { return cx(a.re - b.re, a.im - b.im); }

/* src/fft_simd.rs:220-233: w * z, re = fma(a, x, -(b*y)), im = fma(a, y, b*x) */
static inline oc64 cmul(oc64 w, oc64 z)
{
    double a = w.re, b = w.im, x = z.re, y = z.im;
    return cx(fma(a, x, -b * y), fma(a, y, b * x));
}

/* src/fft_simd.rs:113-120 */
static inline oc64 mulj(int fwd, oc64 z)
{
    return fwd ? cx(-z.im, z.re) : cx(z.im, -z.re);
}

/* src/fft_simd.rs:122-131 */
static inline oc64 mul_e8(int fwd, oc64 z)
{
    oc64 t = cadd(z, mulj(fwd, z));
    return cx(INV_SQRT2 * t.re, INV_SQRT2 * t.im);
}
static inline oc64 mul_ne8(int fwd, oc64 z) { return mul_e8(!fwd, z); }

/* src/fft_simd.rs:133-159 */
static inline oc64 mul_e16(int fwd, oc64 z) { return cmul(cx(H1X, fwd ? H1Y : -H1Y), z); }
static inline oc64 mul_e17(int fwd, oc64 z) { return cmul(cx(-H1Y, fwd ? -H1X : H1X), z); }
static inline oc64 mul_ne16(int fwd, oc64 z) { return mul_e16(!fwd, z); }
static inline oc64 mul_ne17(int fwd, oc64 z) { return mul_e17(!fwd, z); }

/* ------------------------------------------------------------------ */
/* twiddle-free butterflies ("last_butterfly")                         */
/* ------------------------------------------------------------------ */

/* src/dif2.rs:106-113 */
static inline void bf2(oc64 *v)
{
    oc64 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

/* src/dif4.rs:195-214 */
static inline void bf4(int fwd, oc64 *v)
{
    oc64 apc = cadd(v[0], v[2]);
    oc64 amc = csub(v[0], v[2]);
    oc64 bpd = cadd(v[1], v[3]);
    oc64 jbmd = mulj(fwd, csub(v[1], v[3]));
    v[0] = cadd(apc, bpd);
    v[1] = csub(amc, jbmd);
    v[2] = csub(apc, bpd);
    v[3] = cadd(amc, jbmd);
}

/* src/dif8.rs:310-351 (== src/unordered.rs:98-153 without the twiddles) */
static inline void bf8(int fwd, oc64 *v)
{
    oc64 a[4], s[4];
    for (int i = 0; i < 4; i++) {
        a[i] = cadd(v[i], v[i + 4]);
        s[i] = csub(v[i], v[i + 4]);
    }
    oc64 js2 = mulj(fwd, s[2]);
    oc64 js3 = mulj(fwd, s[3]);

    oc64 a02p = cadd(a[0], a[2]);
    oc64 s02m = csub(s[0], js2);
    oc64 a02m = csub(a[0], a[2]);
    oc64 s02p = cadd(s[0], js2);
    oc64 a13p = cadd(a[1], a[3]);
    oc64 w8 = mul_ne8(fwd, csub(s[1], js3));
    oc64 ja13m = mulj(fwd, csub(a[1], a[3]));
    oc64 v8 = mul_e8(fwd, cadd(s[1], js3));

    v[0] = cadd(a02p, a13p);
    v[1] = cadd(s02m, w8);
    v[2] = csub(a02m, ja13m);
    v[3] = csub(s02p, v8);
    v[4] = csub(a02p, a13p);
    v[5] = csub(s02m, w8);
    v[6] = cadd(a02m, ja13m);
    v[7] = cadd(s02p, v8);
}

/* src/dif16.rs:649-772 */
static inline void bf16(int fwd, oc64 *v)
{
    oc64 a[8], s[8];
    for (int i = 0; i < 8; i++) {
        a[i] = cadd(v[i], v[i + 8]);
        s[i] = csub(v[i], v[i + 8]);
    }
    oc64 ap[4], sm[4], am[4], sp[4];
    for (int i = 0; i < 4; i++) {
        oc64 js = mulj(fwd, s[i + 4]);
        ap[i] = cadd(a[i], a[i + 4]);
        sm[i] = csub(s[i], js);
        am[i] = csub(a[i], a[i + 4]);
        sp[i] = cadd(s[i], js);
    }
    /* E = even half (inputs 0,2 groups), O = odd half (inputs 1,3 groups) */
    oc64 t[2][8];
    for (int h = 0; h < 2; h++) {
        int e = h, o = h + 2;
        oc64 w8 = mul_ne8(fwd, sm[o]);
        oc64 j_ = mulj(fwd, am[o]);
        oc64 v8 = mul_e8(fwd, sp[o]);
        t[h][0] = cadd(ap[e], ap[o]);
        t[h][1] = cadd(sm[e], w8);
        t[h][2] = csub(am[e], j_);
        t[h][3] = csub(sp[e], v8);
        t[h][4] = csub(ap[e], ap[o]);
        t[h][5] = csub(sm[e], w8);
        t[h][6] = cadd(am[e], j_);
        t[h][7] = cadd(sp[e], v8);
    }
    const oc64 *E = t[0], *O = t[1];
    oc64 u1 = mul_e16(fwd, O[1]);
    oc64 u2 = mul_ne8(fwd, O[2]);
    oc64 u3 = mul_e17(fwd, O[3]);
    oc64 u4 = mulj(fwd, O[4]);
    oc64 u5 = mul_ne17(fwd, O[5]);
    oc64 u6 = mul_e8(fwd, O[6]);
    oc64 u7 = mul_ne16(fwd, O[7]);

    v[0] = cadd(E[0], O[0]);
    v[1] = cadd(E[1], u1);
    v[2] = cadd(E[2], u2);
    v[3] = cadd(E[3], u3);
    v[4] = csub(E[4], u4);
    v[5] = csub(E[5], u5);
    v[6] = csub(E[6], u6);
    v[7] = csub(E[7], u7);
    v[8] = csub(E[0], O[0]);
    v[9] = csub(E[1], u1);
    v[10] = csub(E[2], u2);
    v[11] = csub(E[3], u3);
    v[12] = cadd(E[4], u4);
    v[13] = cadd(E[5], u5);
    v[14] = cadd(E[6], u6);
    v[15] = cadd(E[7], u7);
}

static inline __attribute__((always_inline)) void bfR(const int R, const int fwd, oc64 *v)
{
    switch (R) {
    case 2: bf2(v); break;
    case 4: bf4(fwd, v); break;
    case 8: bf8(fwd, v); break;
    default: bf16(fwd, v); break;
    }
}

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

/*
 * f128_oracle.c -- CPU restatement of concrete-fft's fft128 negacyclic transform and the
 * double-double ("f128") operations it uses.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Compile with gcc -O2 -ffp-contract=off -mfma.
 *
 * Two multiply variants exist in the reference and round differently in the last bits:
 *   ORC_F128_SCALAR  src/fft128/f128_ops.rs:395-400   e = e + (a0*b1 + a1*b0)
 *   ORC_F128_FMA     src/fft128/f128_ops.rs:837-841   e = fma(a0, b1, fma(a1, b0, e))
 * On x86 hosts with AVX2 the reference dispatches to the FMA form
 * (src/fft128/mod.rs:1041-1071); the GPU kernels follow the FMA form too.
 * The SIMD two_sum (compare-swap + quick_two_sum, f128_ops.rs:635-644) returns the same
 * (s, e) as the branch-free six-operation form used here, because both are error-free.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>

/* src/fft128/f128_ops.rs:6-40 */
static inline void quick_two_sum(double a, double b, double *s, double *e)
{
    double t = a + b;
    *e = b - (t - a);
    *s = t;
}
static inline void two_sum(double a, double b, double *s, double *e)
{
    double t = a + b;
    double bb = t - a;
    *e = (a - (t - bb)) + (b - bb);
    *s = t;
}
static inline void two_diff(double a, double b, double *s, double *e)
{
    double t = a - b;
    double bb = t - a;
    *e = (a - (t - bb)) - (b + bb);
    *s = t;
}
static inline void two_prod(double a, double b, double *p, double *e)
{
    double t = a * b;
    *e = fma(a, b, -t);
    *p = t;
}

static inline of128 dd(double hi, double lo) { of128 r = { hi, lo }; return r; }

of128 orc_f128_add_estimate(of128 a, of128 b)
{
    double s, e;
    two_sum(a.hi, b.hi, &s, &e);
    e = e + (a.lo + b.lo);
    of128 r;
    quick_two_sum(s, e, &r.hi, &r.lo);
    return r;
}

of128 orc_f128_sub_estimate(of128 a, of128 b)
{
    double s, e;
    two_diff(a.hi, b.hi, &s, &e);
    e = e + a.lo;
    e = e - b.lo;
    of128 r;
    quick_two_sum(s, e, &r.hi, &r.lo);
    return r;
}

of128 orc_f128_add(of128 a, of128 b)
{
    double s1, s2, t1, t2;
    two_sum(a.hi, b.hi, &s1, &s2);
    two_sum(a.lo, b.lo, &t1, &t2);
    s2 = s2 + t1;
    quick_two_sum(s1, s2, &s1, &s2);
    s2 = s2 + t2;
    of128 r;
    quick_two_sum(s1, s2, &r.hi, &r.lo);
    return r;
}

of128 orc_f128_sub(of128 a, of128 b)
{
    double s1, s2, t1, t2;
    two_diff(a.hi, b.hi, &s1, &s2);
    two_diff(a.lo, b.lo, &t1, &t2);
    s2 = s2 + t1;
    quick_two_sum(s1, s2, &s1, &s2);
    s2 = s2 + t2;
    of128 r;
    quick_two_sum(s1, s2, &r.hi, &r.lo);
    return r;
}

of128 orc_f128_mul(of128 a, of128 b, int variant)
{
    double p, e;
    two_prod(a.hi, b.hi, &p, &e);
    if (variant == ORC_F128_FMA)
        e = fma(a.hi, b.lo, fma(a.lo, b.hi, e));
    else
        e = e + (a.hi * b.lo + a.lo * b.hi);
    of128 r;
    quick_two_sum(p, e, &r.hi, &r.lo);
    return r;
}

/* f128_ops.rs:286-291 / :380-385 */
static of128 add_f128_f64(of128 a, double b)
{
    double s1, s2;
    two_sum(a.hi, b, &s1, &s2);
    s2 = s2 + a.lo;
    of128 r;
    quick_two_sum(s1, s2, &r.hi, &r.lo);
    return r;
}
static of128 mul_f128_f64(of128 a, double b)
{
    double p1, p2;
    two_prod(a.hi, b, &p1, &p2);
    p2 = p2 + (a.lo * b);
    of128 r;
    quick_two_sum(p1, p2, &r.hi, &r.lo);
    return r;
}

/* f128_ops.rs:477-491 */
of128 orc_f128_div(of128 a, of128 b)
{
    double q1 = a.hi / b.hi;
    of128 r = orc_f128_sub(a, mul_f128_f64(b, q1));
    double q2 = r.hi / b.hi;
    r = orc_f128_sub(r, mul_f128_f64(b, q2));
    double q3 = r.hi / b.hi;
    of128 q;
    quick_two_sum(q1, q2, &q.hi, &q.lo);
    return add_f128_f64(q, q3);
}

/* f128_ops.rs:457-474 */
of128 orc_f128_div_estimate(of128 a, of128 b)
{
    double q1 = a.hi / b.hi;
    of128 r = mul_f128_f64(b, q1);
    double s1, s2;
    two_diff(a.hi, r.hi, &s1, &s2);
    s2 = s2 - r.lo;
    s2 = s2 + a.lo;
    double q2 = (s1 + s2) / b.hi;
    of128 q;
    quick_two_sum(q1, q2, &q.hi, &q.lo);
    return q;
}

/* mixed-operand forms, f128_ops.rs:279-299, 324-347, 373-392, 413-455 */
static of128 sub_f128_f64_(of128 a, double b) /* :331-336 */
{
    double s1, s2;
    two_diff(a.hi, b, &s1, &s2);
    s2 = s2 + a.lo;
    of128 r;
    quick_two_sum(s1, s2, &r.hi, &r.lo);
    return r;
}
static of128 sub_f64_f128_(double a, of128 b) /* :339-345 */
{
    double s1, s2;
    two_diff(a, b.hi, &s1, &s2);
    s2 = s2 - b.lo;
    of128 r;
    quick_two_sum(s1, s2, &r.hi, &r.lo);
    return r;
}
static of128 div_f64_f64_(double a, double b) /* :413-428 */
{
    double q1 = a / b, p1, p2, s, e;
    two_prod(q1, b, &p1, &p2);
    two_diff(a, p1, &s, &e);
    e = e - p2;
    double q2 = (s + e) / b;
    of128 r;
    quick_two_sum(q1, q2, &r.hi, &r.lo);
    return r;
}
static of128 div_f128_f64_(of128 a, double b) /* :431-448 */
{
    double q1 = a.hi / b, p1, p2, s, e;
    two_prod(q1, b, &p1, &p2);
    two_diff(a.hi, p1, &s, &e);
    e = e + a.lo;
    e = e - p2;
    double q2 = (s + e) / b;
    of128 r;
    quick_two_sum(q1, q2, &r.hi, &r.lo);
    return r;
}

void orc_f128_binary_op(int op, const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo,
                        double *out_hi, double *out_lo, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        of128 a = { a_hi[i], a_lo ? a_lo[i] : 0.0 }, b = { b_hi[i], b_lo ? b_lo[i] : 0.0 }, r;
        switch (op) {
        case 0: r = orc_f128_add(a, b); break;
        case 1: r = orc_f128_sub(a, b); break;
        case 2: r = orc_f128_mul(a, b, ORC_F128_SCALAR); break;
        case 3: r = orc_f128_div(a, b); break;
        case 4: r = orc_f128_add_estimate(a, b); break;
        case 5: r = orc_f128_sub_estimate(a, b); break;
        case 6: r = orc_f128_div_estimate(a, b); break;
        case 7: r = add_f128_f64(a, b.hi); break;                        /* add_f128_f64 :286-291 (= add_f64_f128 swapped) */
        case 8: r = sub_f128_f64_(a, b.hi); break;
        case 9: r = sub_f64_f128_(a.hi, b); break;
        case 10: r = mul_f128_f64(a, b.hi); break;                       /* :380-385 (= mul_f64_f128 swapped) */
        case 11: r = div_f128_f64_(a, b.hi); break;
        case 12: r = orc_f128_div(dd(a.hi, 0.0), b); break;              /* div_f64_f128 :451-454 */
        case 13: two_sum(a.hi, b.hi, &r.hi, &r.lo); break;               /* add_f64_f64 :279-283 */
        case 14: two_diff(a.hi, b.hi, &r.hi, &r.lo); break;              /* sub_f64_f64 :324-328 */
        case 15: two_prod(a.hi, b.hi, &r.hi, &r.lo); break;              /* mul_f64_f64 :373-377 */
        default: r = div_f64_f64_(a.hi, b.hi); break;                    /* div_f64_f64 :413-428 */
        }
        out_hi[i] = r.hi;
        out_lo[i] = r.lo;
    }
}

void orc_f128_cplx_mul_scale(double *l_re0, double *l_re1, double *l_im0, double *l_im1, const double *r_re0,
                             const double *r_re1, const double *r_im0, const double *r_im1, double factor, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        of128 ar = { l_re0[i], l_re1[i] }, ai = { l_im0[i], l_im1[i] };
        of128 br = { r_re0[i], r_re1[i] }, bi = { r_im0[i], r_im1[i] };
        of128 rr = orc_f128_mul(ar, br, ORC_F128_SCALAR), ri = orc_f128_mul(ar, bi, ORC_F128_SCALAR);
        of128 ir = orc_f128_mul(ai, br, ORC_F128_SCALAR), ii = orc_f128_mul(ai, bi, ORC_F128_SCALAR);
        of128 pr = orc_f128_sub_estimate(rr, ii), pi = orc_f128_add_estimate(ir, ri);
        l_re0[i] = pr.hi * factor; l_re1[i] = pr.lo * factor;
        l_im0[i] = pi.hi * factor; l_im1[i] = pi.lo * factor;
    }
}

/* helpers used by sincospi only (scalar operator impls, f128_ops.rs:60-230) */
static inline of128 mul_ss(of128 a, of128 b) { return orc_f128_mul(a, b, ORC_F128_SCALAR); }

/* f128_ops.rs:331-336 (sub_f128_f64) */
static of128 sub_f128_f64(of128 a, double b)
{
    double s1, s2;
    two_diff(a.hi, b, &s1, &s2);
    s2 = s2 + a.lo;
    of128 r;
    quick_two_sum(s1, s2, &r.hi, &r.lo);
    return r;
}

of128 orc_f128_sqr(of128 a);
/* f128_ops.rs:404-409 (sqr) */
static of128 sqr_f128(of128 a)
{
    double p1, p2;
    two_prod(a.hi, a.hi, &p1, &p2);
    p2 = p2 + 2.0 * (a.hi * a.lo);
    of128 r;
    quick_two_sum(p1, p2, &r.hi, &r.lo);
    return r;
}

static inline of128 neg_f128(of128 a) { return dd(-a.hi, -a.lo); }

/* f128_ops.rs:578-618 */
static const of128 F128_PI = { 3.141592653589793, 1.2246467991473532e-16 };
static const of128 SINPI_TAYLOR[9] = {
    { -5.16771278004997, 2.2665622825789447e-16 },
    { 2.5501640398773455, -7.931006345326556e-17 },
    { -0.5992645293207921, 2.845026112698218e-17 },
    { 0.08214588661112823, -3.847292805297656e-18 },
    { -0.0073704309457143504, -3.328281165603432e-19 },
    { 0.00046630280576761255, 1.0704561733683463e-20 },
    { -2.1915353447830217e-5, 1.4648526682685598e-21 },
    { 7.952054001475513e-7, 1.736540361519021e-23 },
    { -2.2948428997269873e-8, -7.376346207041088e-26 },
};
static const of128 COSPI_TAYLOR[9] = {
    { -4.934802200544679, -3.1326477543698557e-16 },
    { 4.0587121264167685, -2.6602000824298645e-16 },
    { -1.3352627688545895, 3.1815237892149862e-18 },
    { 0.2353306303588932, -1.2583065576724427e-18 },
    { -0.02580689139001406, 1.170191067939226e-18 },
    { 0.0019295743094039231, -9.669517939986956e-20 },
    { -0.0001046381049248457, -2.421206183964864e-21 },
    { 4.303069587032947e-6, -2.864010082936791e-22 },
    { -1.3878952462213771e-7, -7.479362090417238e-24 },
};
static const of128 SIN_K_PI_16[4] = {
    { 0.19509032201612828, -7.991079068461731e-18 },
    { 0.3826834323650898, -1.0050772696461588e-17 },
    { 0.5555702330196022, 4.709410940561677e-17 },
    { 0.7071067811865476, -4.833646656726457e-17 },
};
static const of128 COS_K_PI_16[4] = {
    { 0.9807852804032304, 1.8546939997825006e-17 },
    { 0.9238795325112867, 1.7645047084336677e-17 },
    { 0.8314696123025452, 1.4073856984728024e-18 },
    { 0.7071067811865476, -4.833646656726457e-17 },
};

/* f128_ops.rs:514-532 */
static void sincospi_taylor(of128 x, of128 *s_out, of128 *c_out)
{
    of128 sinc = F128_PI;
    of128 cosv = dd(1.0, 0.0);
    of128 sq = sqr_f128(x);
    of128 pw = dd(1.0, 0.0);
    for (int i = 0; i < 9; i++) {
        pw = mul_ss(pw, sq);
        sinc = orc_f128_add(sinc, mul_ss(SINPI_TAYLOR[i], pw));
        cosv = orc_f128_add(cosv, mul_ss(COSPI_TAYLOR[i], pw));
    }
    *s_out = mul_ss(sinc, x);
    *c_out = cosv;
}

/* f128_ops.rs:534-575; input in [-1, 1] */
void orc_f128_sincospi(of128 x, of128 *s_out, of128 *c_out)
{
    double p = round(x.hi * 2.0);
    of128 r = sub_f128_f64(x, p * 0.5);
    double q = round(r.hi * 16.0);
    r = sub_f128_f64(r, q * (1.0 / 16.0));

    long pi = (long)p, qi = (long)q;
    unsigned long q_abs = (unsigned long)(qi < 0 ? -qi : qi);

    of128 sin_r, cos_r, s, c;
    sincospi_taylor(r, &sin_r, &cos_r);
    if (qi == 0) {
        s = sin_r;
        c = cos_r;
    } else {
        of128 u = COS_K_PI_16[q_abs - 1];
        of128 v = SIN_K_PI_16[q_abs - 1];
        if (qi > 0) {
            s = orc_f128_add(mul_ss(u, sin_r), mul_ss(v, cos_r));
            c = orc_f128_sub(mul_ss(u, cos_r), mul_ss(v, sin_r));
        } else {
            s = orc_f128_sub(mul_ss(u, sin_r), mul_ss(v, cos_r));
            c = orc_f128_add(mul_ss(u, cos_r), mul_ss(v, sin_r));
        }
    }
    if (pi == 0) {
        *s_out = s; *c_out = c;
    } else if (pi == 1) {
        *s_out = c; *c_out = neg_f128(s);
    } else if (pi == -1) {
        *s_out = neg_f128(c); *c_out = s;
    } else {
        *s_out = neg_f128(s); *c_out = neg_f128(c);
    }
}

/* src/fft128/mod.rs:1794-1803 */
static size_t bitreverse(size_t i, size_t n)
{
    unsigned logn = 0;
    while (((size_t)1 << logn) < n) logn++;
    size_t r = 0;
    for (unsigned k = 0; k < logn; k++)
        r |= ((i >> k) & 1) << (logn - k - 1);
    return r;
}

/* src/fft128/mod.rs:1805-1828 */
void orc_f128_init_twiddles(size_t n, double *re0, double *re1, double *im0, double *im1)
{
    for (size_t m = 1; m < n; m *= 2) {
        for (size_t i = 0; i < m; i++) {
            size_t k = 2 * m + i, pos = m + i;
            of128 th = dd((double)bitreverse(k, 2 * n) / (double)(2 * n), 0.0);
            of128 s, c;
            orc_f128_sincospi(th, &s, &c);
            re0[pos] = c.hi; re1[pos] = c.lo;
            im0[pos] = s.hi; im1[pos] = s.lo;
        }
    }
}

struct orc_f128_plan {
    size_t n;
    double *tw[4];
};

orc_f128_plan *orc_f128_plan_new(size_t n)
{
    if (n < 32 || (n & (n - 1)))
        return NULL; /* src/fft128/mod.rs:1865-1866 */
    orc_f128_plan *p = calloc(1, sizeof *p);
    p->n = n;
    for (int i = 0; i < 4; i++)
        p->tw[i] = calloc(n, sizeof(double));
    orc_f128_init_twiddles(n, p->tw[0], p->tw[1], p->tw[2], p->tw[3]);
    return p;
}

void orc_f128_plan_free(orc_f128_plan *p)
{
    if (!p) return;
    for (int i = 0; i < 4; i++) free(p->tw[i]);
    free(p);
}

const double *orc_f128_twiddles(const orc_f128_plan *p, int which) { return p->tw[which & 3]; }


/* Inner loops of one butterfly block written on plain doubles so that gcc -O3 vectorises them (the
 * reference's AVX2 path, src/fft128/mod.rs:406-662, does the same work four lanes at a time).  Same
 * operations in the same order as orc_f128_{add,sub}_estimate / orc_f128_mul(FMA): same bits. */
#define DD_TWO_SUM(a, b, s, e) do { double t_ = (a) + (b); double bb_ = t_ - (a); (e) = ((a) - (t_ - bb_)) + ((b) - bb_); (s) = t_; } while (0)
#define DD_TWO_DIFF(a, b, s, e) do { double t_ = (a) - (b); double bb_ = t_ - (a); (e) = ((a) - (t_ - bb_)) - ((b) + bb_); (s) = t_; } while (0)
#define DD_QTS(a, b, s, e) do { double t_ = (a) + (b); (e) = (b) - (t_ - (a)); (s) = t_; } while (0)
#define DD_ADD(ah, al, bh, bl, rh, rl) do { double s_, e_; DD_TWO_SUM(ah, bh, s_, e_); e_ = e_ + ((al) + (bl)); DD_QTS(s_, e_, rh, rl); } while (0)
#define DD_SUB(ah, al, bh, bl, rh, rl) do { double s_, e_; DD_TWO_DIFF(ah, bh, s_, e_); e_ = e_ + (al); e_ = e_ - (bl); DD_QTS(s_, e_, rh, rl); } while (0)
#define DD_MUL(ah, al, bh, bl, rh, rl) do { double p_ = (ah) * (bh); double e_ = fma(ah, bh, -p_); e_ = fma(ah, bl, fma(al, bh, e_)); DD_QTS(p_, e_, rh, rl); } while (0)

static void fwd_block_fma(double *restrict re0, double *restrict re1, double *restrict im0, double *restrict im1,
                          size_t start, size_t t, double wrh, double wrl, double wih, double wil)
{
    for (size_t j = start; j < start + t; j++) {
        double z1rh = re0[j + t], z1rl = re1[j + t], z1ih = im0[j + t], z1il = im1[j + t];
        double rrh, rrl, rih, ril, irh, irl, iih, iil, zwrh, zwrl, zwih, zwil;
        DD_MUL(z1rh, z1rl, wrh, wrl, rrh, rrl);
        DD_MUL(z1rh, z1rl, wih, wil, rih, ril);
        DD_MUL(z1ih, z1il, wrh, wrl, irh, irl);
        DD_MUL(z1ih, z1il, wih, wil, iih, iil);
        DD_SUB(rrh, rrl, iih, iil, zwrh, zwrl);
        DD_ADD(irh, irl, rih, ril, zwih, zwil);
        double z0rh = re0[j], z0rl = re1[j], z0ih = im0[j], z0il = im1[j];
        double a, b;
        DD_ADD(z0rh, z0rl, zwrh, zwrl, a, b); re0[j] = a; re1[j] = b;
        DD_ADD(z0ih, z0il, zwih, zwil, a, b); im0[j] = a; im1[j] = b;
        DD_SUB(z0rh, z0rl, zwrh, zwrl, a, b); re0[j + t] = a; re1[j + t] = b;
        DD_SUB(z0ih, z0il, zwih, zwil, a, b); im0[j + t] = a; im1[j + t] = b;
    }
}

static void inv_block_fma(double *restrict re0, double *restrict re1, double *restrict im0, double *restrict im1,
                          size_t start, size_t t, double wrh, double wrl, double wih, double wil)
{
    for (size_t j = start; j < start + t; j++) {
        double z0rh = re0[j], z0rl = re1[j], z0ih = im0[j], z0il = im1[j];
        double z1rh = re0[j + t], z1rl = re1[j + t], z1ih = im0[j + t], z1il = im1[j + t];
        double drh, drl, dih, dil, a, b;
        DD_SUB(z0rh, z0rl, z1rh, z1rl, drh, drl);
        DD_SUB(z0ih, z0il, z1ih, z1il, dih, dil);
        DD_ADD(z0rh, z0rl, z1rh, z1rl, a, b); re0[j] = a; re1[j] = b;
        DD_ADD(z0ih, z0il, z1ih, z1il, a, b); im0[j] = a; im1[j] = b;
        double rrh, rrl, rih, ril, irh, irl, iih, iil;
        DD_MUL(drh, drl, wrh, wrl, rrh, rrl);
        DD_MUL(drh, drl, wih, wil, rih, ril);
        DD_MUL(dih, dil, wrh, wrl, irh, irl);
        DD_MUL(dih, dil, wih, wil, iih, iil);
        DD_ADD(rrh, rrl, iih, iil, a, b); re0[j + t] = a; re1[j + t] = b;
        DD_SUB(irh, irl, rih, ril, a, b); im0[j + t] = a; im1[j + t] = b;
    }
}

/* src/fft128/mod.rs:352-402 (cplx_mul :310-326) */
void orc_f128_fwd(const orc_f128_plan *p, double *re0, double *re1, double *im0, double *im1, int variant)
{
    size_t n = p->n, t = n;
    for (size_t m = 1; m < n; m *= 2) {
        t /= 2;
        for (size_t i = 0; i < m; i++) {
            of128 wr = dd(p->tw[0][m + i], p->tw[1][m + i]);
            of128 wi = dd(p->tw[2][m + i], p->tw[3][m + i]);
            size_t start = 2 * i * t;
            if (variant == ORC_F128_FMA) {
                fwd_block_fma(re0, re1, im0, im1, start, t, wr.hi, wr.lo, wi.hi, wi.lo);
                continue;
            }
            for (size_t j = start; j < start + t; j++) {
                of128 z0r = dd(re0[j], re1[j]), z0i = dd(im0[j], im1[j]);
                of128 z1r = dd(re0[j + t], re1[j + t]), z1i = dd(im0[j + t], im1[j + t]);
                of128 rr = orc_f128_mul(z1r, wr, variant);
                of128 ri = orc_f128_mul(z1r, wi, variant);
                of128 ir = orc_f128_mul(z1i, wr, variant);
                of128 ii = orc_f128_mul(z1i, wi, variant);
                of128 zwr = orc_f128_sub_estimate(rr, ii);
                of128 zwi = orc_f128_add_estimate(ir, ri);
                of128 o0r = orc_f128_add_estimate(z0r, zwr), o0i = orc_f128_add_estimate(z0i, zwi);
                of128 o1r = orc_f128_sub_estimate(z0r, zwr), o1i = orc_f128_sub_estimate(z0i, zwi);
                re0[j] = o0r.hi; re1[j] = o0r.lo; im0[j] = o0i.hi; im1[j] = o0i.lo;
                re0[j + t] = o1r.hi; re1[j + t] = o1r.lo; im0[j + t] = o1i.hi; im1[j + t] = o1i.lo;
            }
        }
    }
}

/* src/fft128/mod.rs:1105-1155 (cplx_mul_conj :330-346) */
void orc_f128_inv(const orc_f128_plan *p, double *re0, double *re1, double *im0, double *im1, int variant)
{
    size_t n = p->n, t = 1, m = n;
    while (m > 1) {
        m /= 2;
        for (size_t i = 0; i < m; i++) {
            of128 wr = dd(p->tw[0][m + i], p->tw[1][m + i]);
            of128 wi = dd(p->tw[2][m + i], p->tw[3][m + i]);
            size_t start = 2 * i * t;
            if (variant == ORC_F128_FMA) {
                inv_block_fma(re0, re1, im0, im1, start, t, wr.hi, wr.lo, wi.hi, wi.lo);
                continue;
            }
            for (size_t j = start; j < start + t; j++) {
                of128 z0r = dd(re0[j], re1[j]), z0i = dd(im0[j], im1[j]);
                of128 z1r = dd(re0[j + t], re1[j + t]), z1i = dd(im0[j + t], im1[j + t]);
                of128 dr = orc_f128_sub_estimate(z0r, z1r), di = orc_f128_sub_estimate(z0i, z1i);
                of128 o0r = orc_f128_add_estimate(z0r, z1r), o0i = orc_f128_add_estimate(z0i, z1i);
                of128 rr = orc_f128_mul(dr, wr, variant);
                of128 ri = orc_f128_mul(dr, wi, variant);
                of128 ir = orc_f128_mul(di, wr, variant);
                of128 ii = orc_f128_mul(di, wi, variant);
                of128 o1r = orc_f128_add_estimate(rr, ii);
                of128 o1i = orc_f128_sub_estimate(ir, ri);
                re0[j] = o0r.hi; re1[j] = o0r.lo; im0[j] = o0i.hi; im1[j] = o0i.lo;
                re0[j + t] = o1r.hi; re1[j + t] = o1r.lo; im0[j + t] = o1i.hi; im1[j + t] = o1i.lo;
            }
        }
        t *= 2;
    }
}

struct fbatch { const orc_f128_plan *p; double *a[4]; int inverse, variant; };
static void fbatch_rows(void *ctx, size_t lo, size_t hi)
{
    struct fbatch *f = ctx;
    for (size_t b = lo; b < hi; b++) {
        size_t o = b * f->p->n;
        if (f->inverse)
            orc_f128_inv(f->p, f->a[0] + o, f->a[1] + o, f->a[2] + o, f->a[3] + o, f->variant);
        else
            orc_f128_fwd(f->p, f->a[0] + o, f->a[1] + o, f->a[2] + o, f->a[3] + o, f->variant);
    }
}

void orc_f128_fwd_batch(const orc_f128_plan *p, double *re0, double *re1, double *im0, double *im1,
                        size_t batch, int variant, int threads)
{
    struct fbatch f = { p, { re0, re1, im0, im1 }, 0, variant };
    orc_parallel_rows(threads, batch, fbatch_rows, &f);
}

void orc_f128_inv_batch(const orc_f128_plan *p, double *re0, double *re1, double *im0, double *im1,
                        size_t batch, int variant, int threads)
{
    struct fbatch f = { p, { re0, re1, im0, im1 }, 1, variant };
    orc_parallel_rows(threads, batch, fbatch_rows, &f);
}

of128 orc_f128_sqr(of128 a) { return sqr_f128(a); }

/* unary operators on arrays; op codes match include/cfft_b200.h: 0 sqr :404-409, 1 abs :506-511, 2 neg (Neg impl :232-238),
 * 3 sincospi :534-575 (out = sin, out2 = cos; inputs must lie in [-1, 1]), 4 is_nan :499-501 (out_hi = 1.0 / 0.0, out_lo = 0) */
void orc_f128_unary_op(int op, const double *a_hi, const double *a_lo, double *out_hi, double *out_lo, double *out2_hi,
                       double *out2_lo, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        of128 a = { a_hi[i], a_lo[i] }, r = { 0.0, 0.0 }, r2 = { 0.0, 0.0 };
        switch (op) {
        case 0: r = sqr_f128(a); break;
        case 1: r = a.hi < 0.0 ? neg_f128(a) : a; break;
        case 2: r = neg_f128(a); break;
        case 3: orc_f128_sincospi(a, &r, &r2); break;
        default: r.hi = (a.hi != a.hi || a.lo != a.lo) ? 1.0 : 0.0; break;
        }
        out_hi[i] = r.hi;
        out_lo[i] = r.lo;
        if (op == 3) {
            out2_hi[i] = r2.hi;
            out2_lo[i] = r2.lo;
        }
    }
}

/* PartialOrd / PartialEq, f128_ops.rs:240-274: -1 Less, 0 Equal, 1 Greater, 2 None (unordered); b_lo == NULL: an f64 operand */
static int cmp_f64(double x, double y) { return x < y ? -1 : (x > y ? 1 : (x == y ? 0 : 2)); }
void orc_f128_compare(const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo, signed char *out, size_t len)
{
    for (size_t i = 0; i < len; i++) {
        const int first = cmp_f64(a_hi[i], b_hi[i]);
        out[i] = (signed char)(first == 0 ? cmp_f64(a_lo[i], b_lo ? b_lo[i] : 0.0) : first);
    }
}

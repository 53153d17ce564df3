/*
 * oracle.h -- CPU restatement of the concrete-fft v0.5.1 transform hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (concrete_fft_b200/,
 * include/, the C-ABI library) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / reported CPU baseline.
 *
 * Parity pin: the c64 part reproduces the reference's own 2048-entry golden
 * vector (src/unordered.rs:1176-9396) bit for bit (tests/test_oracle_golden.py).
 * The fft128 part has no known-answer vector in the reference;
 * it is pinned by the reference's negacyclic-convolution property
 * (src/fft128/mod.rs:1972-2065) and the f128 op error bounds -- "bit-level
 * parity unpinned" for fft128.
 *
 * All citations are file:line under /root/reference.
 */
#ifndef CFFT_ORACLE_H
#define CFFT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } oc64; /* src/lib.rs:84 (num_complex::Complex64) */

/* FftAlgo discriminants, src/ordered.rs:28-45 */
enum { ORC_DIF2 = 0, ORC_DIT2, ORC_DIF4, ORC_DIT4, ORC_DIF8, ORC_DIT8, ORC_DIF16, ORC_DIT16 };

/* src/fft_simd.rs:237-296 */
void orc_sincospi64(double a, double *s, double *c);

/* src/fft_simd.rs:298-321: w, w_inv have 2n entries each (NaN filled first). */
void orc_init_wt(size_t r, size_t n, oc64 *w, oc64 *w_inv);

/* ordered::Plan (src/ordered.rs:187-374).  n = 2^k, k <= 10. Returns NULL on bad args. */
typedef struct orc_ordered_plan orc_ordered_plan;
orc_ordered_plan *orc_ordered_plan_new(size_t n, int algo);
void orc_ordered_plan_free(orc_ordered_plan *);
void orc_ordered_fwd(const orc_ordered_plan *, oc64 *buf, oc64 *scratch /* n */);
void orc_ordered_inv(const orc_ordered_plan *, oc64 *buf, oc64 *scratch /* n */);

/* unordered::Plan (src/unordered.rs:496-940), Method::UserProvided. */
typedef struct orc_unordered_plan orc_unordered_plan;
orc_unordered_plan *orc_unordered_plan_new(size_t n, int base_algo, size_t base_n);
void orc_unordered_plan_free(orc_unordered_plan *);
void orc_unordered_fwd(const orc_unordered_plan *, oc64 *buf, oc64 *scratch /* base_n */);
void orc_unordered_inv(const orc_unordered_plan *, oc64 *buf, oc64 *scratch /* base_n */);
void orc_unordered_fwd_monomial(const orc_unordered_plan *, size_t degree, oc64 *buf);
/* raw twiddle tables (n + base_n entries), scalar layout (complex_per_reg = 1) */
const oc64 *orc_unordered_twiddles(const orc_unordered_plan *, int inverse);

/* split [0, total) into `threads` contiguous ranges, one pthread each (threads <= 1: inline) */
void orc_parallel_rows(int threads, size_t total, void (*fn)(void *ctx, size_t lo, size_t hi), void *ctx);

/* batch helpers: one independent transform per row of n entries */
void orc_unordered_fwd_batch(const orc_unordered_plan *, oc64 *buf, size_t batch, int threads);
void orc_unordered_inv_batch(const orc_unordered_plan *, oc64 *buf, size_t batch, int threads);

/* ---- caller-side steps around the c64 transform (poly_oracle.c; semantics fixed by this library, see its header) ---- */
void orc_poly_twist_tables(size_t n, oc64 *twist, oc64 *untwist);
void orc_poly_fold_twist(size_t n, int torus, const int64_t *poly, const oc64 *twist, oc64 *out);
void orc_poly_untwist_round(size_t n, int torus, int accumulate, const oc64 *z, const oc64 *untwist, int64_t *out);
void orc_poly_mul_batch(const orc_unordered_plan *plan, size_t n, size_t base_n, const int64_t *a, size_t k_terms, const oc64 *b,
                        size_t b_row_stride, int64_t *out, size_t batch, int torus, int accumulate, int threads);

/* src/unordered.rs:1039-1059 */
size_t orc_bit_rev(unsigned nbits, size_t i);
size_t orc_bit_rev_twice(unsigned nbits, unsigned base_nbits, size_t i);
size_t orc_bit_rev_twice_inv(unsigned nbits, unsigned base_nbits, size_t i);

/* ---- fft128 (src/fft128/mod.rs, src/fft128/f128_ops.rs) ---- */
typedef struct { double hi, lo; } of128; /* src/fft128/mod.rs:3-7 */

/* variant: 0 = scalar mul (f128_ops.rs:395-400), 1 = FMA mul (f128_ops.rs:837-841,
 * what the AVX2/AVX-512 paths and the GPU kernel compute). */
enum { ORC_F128_SCALAR = 0, ORC_F128_FMA = 1 };

of128 orc_f128_add_estimate(of128 a, of128 b); /* f128_ops.rs:302-307 */
of128 orc_f128_sub_estimate(of128 a, of128 b); /* f128_ops.rs:350-356 */
of128 orc_f128_add(of128 a, of128 b);          /* f128_ops.rs:311-321 */
of128 orc_f128_sub(of128 a, of128 b);          /* f128_ops.rs:360-370 */
of128 orc_f128_mul(of128 a, of128 b, int variant);
void orc_f128_sincospi(of128 x, of128 *s, of128 *c); /* f128_ops.rs:514-575 */
of128 orc_f128_div(of128 a, of128 b);          /* f128_ops.rs:477-491 */
of128 orc_f128_div_estimate(of128 a, of128 b); /* f128_ops.rs:457-474 */
of128 orc_f128_sqr(of128 a);                   /* f128_ops.rs:404-409 */

/* element-wise array forms of the scalar f128 operators; op codes match include/cfft_b200.h:
 * 0 add, 1 sub, 2 mul (scalar form :395-400), 3 div, 4 add_estimate, 5 sub_estimate, 6 div_estimate */
void orc_f128_binary_op(int op, const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo,
                        double *out_hi, double *out_lo, size_t len);
/* 7 add_f128_f64, 8 sub_f128_f64, 9 sub_f64_f128, 10 mul_f128_f64, 11 div_f128_f64, 12 div_f64_f128, 13 add_f64_f64,
 * 14 sub_f64_f64, 15 mul_f64_f64, 16 div_f64_f64 (f64 operands: the lo pointer may be NULL) */
void orc_f128_unary_op(int op, const double *a_hi, const double *a_lo, double *out_hi, double *out_lo, double *out2_hi,
                       double *out2_lo, size_t len);
void orc_f128_compare(const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo, signed char *out, size_t len);
/* lhs <- (lhs * rhs) * factor on planar double-double complex arrays, scalar cplx_mul
 * (src/fft128/mod.rs:310-326) then four f64 scalings, exactly the loop at src/fft128/mod.rs:2033-2047 */
/* element-wise c64 products with num_complex's `*` / `+` semantics (no FMA): lhs <- lhs * rhs, or
 * acc <- acc + a * b when acc != NULL (the caller-side Fourier-domain step, README.md:10-17) */
void orc_c64_pointwise(double *acc, double *a, const double *b, size_t len);
void orc_f128_cplx_mul_scale(double *l_re0, double *l_re1, double *l_im0, double *l_im1, const double *r_re0,
                             const double *r_re1, const double *r_im0, const double *r_im1, double factor, size_t len);

/* src/fft128/mod.rs:1805-1828: four arrays of n doubles, entry 0 untouched (0.0). */
void orc_f128_init_twiddles(size_t n, double *re0, double *re1, double *im0, double *im1);

typedef struct orc_f128_plan orc_f128_plan;
orc_f128_plan *orc_f128_plan_new(size_t n); /* n = 2^k >= 32 */
void orc_f128_plan_free(orc_f128_plan *);
void orc_f128_fwd(const orc_f128_plan *, double *re0, double *re1, double *im0, double *im1, int variant);
void orc_f128_inv(const orc_f128_plan *, double *re0, double *re1, double *im0, double *im1, int variant);
void orc_f128_fwd_batch(const orc_f128_plan *, double *re0, double *re1, double *im0, double *im1,
                        size_t batch, int variant, int threads);
void orc_f128_inv_batch(const orc_f128_plan *, double *re0, double *re1, double *im0, double *im1,
                        size_t batch, int variant, int threads);
const double *orc_f128_twiddles(const orc_f128_plan *, int which /* 0 re0,1 re1,2 im0,3 im1 */);

#ifdef __cplusplus
}
#endif
#endif

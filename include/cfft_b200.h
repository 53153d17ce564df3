/*
 * cfft_b200.h -- C ABI of the B200-native batched FFT that replaces concrete-fft's
 * transform hot path (zama-ai/concrete-fft v0.5.1).
 *
 * The reference has no FFI layer: its boundary is the Rust public API
 * (ordered::Plan, unordered::Plan, fft128::Plan).  Each entry point below names the Rust
 * item it stands in for (file:line under the reference tree); a thin Rust crate
 * (rust/concrete-fft-b200) keeps those Rust signatures and forwards
 * to these symbols.
 *
 * Conventions
 *  - Plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL =
 *    the legacy default stream).
 *  - Every function returns a cfft_status (0 = ok, < 0 = error) and never aborts; the Rust
 *    shim turns a non-zero status into the panic the reference would raise.
 *  - Plans are immutable after creation: fwd/inv may be called concurrently from many host
 *    threads on one plan (the guarantee `&self` gives in the reference).
 *  - Batched: `batch` independent transforms stored back to back (row stride = fft size).
 *    The reference API is one polynomial per call, i.e. batch = 1.
 *  - Transforms are unnormalised: inv(fwd(x)) = n * x (src/ordered.rs:7-11).
 *  - c64 = { double re, im }, 16 bytes (src/lib.rs:84).
 *  - There is no CPU fallback: every compute entry point needs a CUDA device and fails with
 *    CFFT_ECUDA otherwise.
 */
#ifndef CFFT_B200_H
#define CFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t cfft_status;
enum {
    CFFT_OK = 0,
    CFFT_EINVAL = -1,       /* a precondition the reference asserts on was violated */
    CFFT_ECUDA = -2,        /* CUDA runtime error, see cfft_last_error() */
    CFFT_ENOMEM = -3,
    CFFT_EUNSUPPORTED = -4, /* valid in the reference, not implemented here (none at present) */
    CFFT_ELENGTH = -5       /* buffer length != fft size (assert_eq! at src/unordered.rs:827) */
};

/* ordered::FftAlgo, src/ordered.rs:28-45 (same discriminant order) */
enum {
    CFFT_DIF2 = 0, CFFT_DIT2, CFFT_DIF4, CFFT_DIT4, CFFT_DIF8, CFFT_DIT8, CFFT_DIF16, CFFT_DIT16
};

/* ordered::Method / unordered::Method, src/ordered.rs:50-59, src/unordered.rs:526-537 */
enum {
    CFFT_METHOD_USER = 0,   /* Method::UserProvided */
    CFFT_METHOD_MEASURE = 1 /* Method::Measure(_): on-device autotune instead of CPU timing */
};

typedef struct cfft_plan cfft_plan;

/* ---- plan lifetime ---------------------------------------------------------------- */

/* ordered::Plan::new(n, method), src/ordered.rs:242-278.
 * n must be a power of two.  The reference also asserts n <= 2^10 (src/ordered.rs:244);
 * `allow_large` != 0 lifts that cap up to 2^20 (an extension: BASELINE.json config 3 asks
 * for a standard-order N = 2^16 transform the reference cannot build). */
cfft_status cfft_ordered_plan_create(cfft_plan **out, int device, uint64_t n, int method,
                                     int algo, int allow_large);

/* unordered::Plan::new(n, method), src/unordered.rs:659-747.
 * USER: base_n power of two, <= n, <= 1024, >= 32 unless == n (src/unordered.rs:664-669).
 * MEASURE: base_algo / base_n are ignored; the choice is made on the device and is a pure
 * function of (n, device kind), never of a timing race across runs. */
cfft_status cfft_unordered_plan_create(cfft_plan **out, int device, uint64_t n, int method,
                                       int base_algo, uint64_t base_n);

/* fft128::Plan::new(n), src/fft128/mod.rs:1864-1881: n power of two, >= 32. */
cfft_status cfft_f128_plan_create(cfft_plan **out, int device, uint64_t n);

/* Drop for the three Plan types. */
void cfft_plan_destroy(cfft_plan *plan);

/* Clone for the three Plan types (#[derive(Clone)], src/unordered.rs:495). */
cfft_status cfft_plan_clone(const cfft_plan *plan, cfft_plan **out);
/* The same with the replica's tables on another GPU of the box: same transform, same Fourier-domain order, same tuned
 * kernel variant.  Replicas are what cfft_c64_host_multi / cfft_f128_host_multi shard one host call over. */
cfft_status cfft_plan_clone_to_device(const cfft_plan *plan, int device, cfft_plan **out);

/* ---- plan queries ----------------------------------------------------------------- */

/* Plan::fft_size, src/ordered.rs:291-293, src/unordered.rs:760-762, src/fft128/mod.rs:1891-1893 */
uint64_t cfft_plan_fft_size(const cfft_plan *plan);

/* ordered::Plan::algo (src/ordered.rs:305-307) / unordered::Plan::algo (src/unordered.rs:783-785).
 * For an ordered plan *base_n = n.  CFFT_EINVAL for an fft128 plan. */
cfft_status cfft_plan_algo(const cfft_plan *plan, int *algo, uint64_t *base_n);

/* Plan::fft_scratch, src/ordered.rs:320-322 (n c64), src/unordered.rs:798-800 (base_n c64).
 * Size in bytes / alignment the reference's StackReq would carry.  The device path needs no
 * caller scratch; the Rust shim reports this so that callers sizing a PodStack keep working. */
cfft_status cfft_plan_scratch_req(const cfft_plan *plan, uint64_t *bytes, uint64_t *align);

/* 0 = ordered, 1 = unordered, 2 = fft128 */
int cfft_plan_kind(const cfft_plan *plan);
int cfft_plan_device(const cfft_plan *plan);

/* Which kernel family serves this plan: "fast-b256-regs", "fast-b256-cluster", "fast-b256-column+rows",
 * "fast-b256-column+fused-rows", "fast-b256-persistent-2pass", "ordered-b256-regs-std",
 * "ordered-b256-column+rows-std", "ord16-regs", "exact-regs-spec" (stage schedule built at compile time), "exact-regs",
 * "exact-tile", "f128-radix8-tile", with
 * "/L2-chunked" appended when the passes run chunk by chunk. */
const char *cfft_plan_kernel_name(const cfft_plan *plan);

/* ---- on-device autotune ------------------------------------------------------------ */

/* Replaces measure_fastest (src/ordered.rs:99-180, src/unordered.rs:553-640): times every kernel
 * VARIANT available for this plan (fused vs multi-pass kernels, tile sizes) on `batch_hint`
 * synthetic transforms with CUDA events and keeps the fastest (batch_hint = 0: 128 MiB worth).
 * It never changes (base_algo, base_n) -- the Fourier-domain order -- nor any output bit: all
 * variants of a plan are bit-identical.  Runs implicitly when a plan is created with
 * CFFT_METHOD_MEASURE (skip with env CFFT_B200_NO_AUTOTUNE=1).  Must not race with transforms on
 * the same plan. */
cfft_status cfft_plan_autotune(cfft_plan *plan, uint64_t batch_hint);
/* "variant-name: ms" lines of the last autotune ("" if none); returns bytes written (excl. NUL) */
uint64_t cfft_plan_tuning_report(const cfft_plan *plan, char *buf, uint64_t buf_len);

/* ---- c64 transforms --------------------------------------------------------------- */

/* {ordered,unordered}::Plan::fwd / inv on device memory, in place, stream ordered.
 * src/ordered.rs:342-373, src/unordered.rs:826-839, 927-940.
 * dev_buf: batch * n c64 on the plan's device, 16-byte aligned (CFFT_EINVAL otherwise; host
 * entry points accept any c64 alignment).  Unordered: fwd output / inv input are in the
 * plan's permuted order, index for index as the reference (src/unordered.rs:1046-1051). */
cfft_status cfft_c64_fwd(const cfft_plan *plan, void *dev_buf, uint64_t batch, void *stream);
cfft_status cfft_c64_inv(const cfft_plan *plan, void *dev_buf, uint64_t batch, void *stream);

/* The same with rows `row_stride` >= n c64 apart (row r starts at dev_buf + r * row_stride): the stride_elems of
 * SURVEY.md section 8b, for callers whose polynomials sit inside larger records (Plan::fwd on buf[r * stride ..][..n]
 * for every r).  Elements between the rows are never touched.  row_stride < n: CFFT_EINVAL.  Plans with a single
 * fused kernel (n = 256 .. 8192, base (Dif16, 256), and the ordered extension 2^11 .. 2^13) read and write the
 * strided rows directly; other plans go through a packed workspace (same bits, two extra passes over the data). */
cfft_status cfft_c64_fwd_strided(const cfft_plan *plan, void *dev_buf, uint64_t row_stride, uint64_t batch, void *stream);
cfft_status cfft_c64_inv_strided(const cfft_plan *plan, void *dev_buf, uint64_t row_stride, uint64_t batch, void *stream);

/* Same on HOST memory, synchronous: H2D, transform, D2H through an internal pinned, chunked,
 * double-buffered pipeline.  This is the literal drop-in for Plan::fwd(&mut [c64], stack)
 * (batch = 1) and the entry bench.py times as `e2e`.  `len` = number of c64 in host_buf and
 * must equal batch * n (CFFT_ELENGTH otherwise: the assert_eq! of src/unordered.rs:827). */
cfft_status cfft_c64_fwd_host(const cfft_plan *plan, void *host_buf, uint64_t len, uint64_t batch);
cfft_status cfft_c64_inv_host(const cfft_plan *plan, void *host_buf, uint64_t len, uint64_t batch);
/* fwd immediately followed by inv on the device between one H2D and one D2H (the
 * BASELINE.json "fwd+inv" step); result = n * input. */
cfft_status cfft_c64_fwd_inv_host(const cfft_plan *plan, void *host_buf, uint64_t len, uint64_t batch);

/* One host call sharded over several GPUs (north_star item 5 inside the library): `plans` = nplans replicas of one plan
 * (cfft_plan_clone_to_device), the batch is cut into nplans contiguous row ranges, each range runs through its
 * replica's own H2D / kernels / D2H pipeline on its own host thread; no collective, no peer traffic.  op: 0 fwd, 1 inv,
 * 2 fwd then inv.  Same results, bit for bit, as the single-GPU entry points. */
cfft_status cfft_c64_host_multi(const cfft_plan *const *plans, int nplans, int op, void *host_buf, uint64_t len, uint64_t batch);

/* unordered::Plan::fwd_monomial(degree, buf), src/unordered.rs:844-900: writes the permuted
 * forward transform of X^degree.  degree < n (CFFT_EINVAL otherwise). */
cfft_status cfft_unordered_fwd_monomial(const cfft_plan *plan, uint64_t degree, void *dev_buf,
                                        void *stream);
cfft_status cfft_unordered_fwd_monomial_host(const cfft_plan *plan, uint64_t degree,
                                             void *host_buf, uint64_t len);

/* ---- standard-order mapping (serde) ------------------------------------------------ */

/* bit_rev_twice table: out[i] = position of Fourier coefficient i in the plan's buffer,
 * src/unordered.rs:1046-1051.  `out` has n entries (host).  Identity for ordered plans. */
cfft_status cfft_unordered_permutation(const cfft_plan *plan, uint64_t *out);

/* Device gather / scatter behind serialize_fourier_buffer / deserialize_fourier_buffer,
 * src/unordered.rs:951-1036: dst[b][i] = src[b][perm[i]]  /  dst[b][perm[i]] = src[b][i].
 * src and dst must not overlap. */
cfft_status cfft_unordered_to_standard(const cfft_plan *plan, const void *dev_src, void *dev_dst,
                                       uint64_t batch, void *stream);
cfft_status cfft_unordered_from_standard(const cfft_plan *plan, const void *dev_src, void *dev_dst,
                                         uint64_t batch, void *stream);
/* Host versions (the exact loops of src/unordered.rs:967-969 and :1019-1025 on host memory).
 * from_standard returns CFFT_ELENGTH when count != n (serde invalid_length, :1027-1028). */
cfft_status cfft_unordered_to_standard_host(const cfft_plan *plan, const void *src, void *dst);
cfft_status cfft_unordered_from_standard_host(const cfft_plan *plan, const void *src,
                                              uint64_t count, void *dst);

/* ---- fft128 ------------------------------------------------------------------------ */

/* fft128::Plan::fwd / inv, src/fft128/mod.rs:1905-1960.  Four planar arrays of batch * n
 * doubles (re hi, re lo, im hi, im lo), in place.  Output of fwd / input of inv is in
 * bit-reversed order exactly as the reference. */
cfft_status cfft_f128_fwd(const cfft_plan *plan, double *re0, double *re1, double *im0,
                          double *im1, uint64_t batch, void *stream);
cfft_status cfft_f128_inv(const cfft_plan *plan, double *re0, double *re1, double *im0,
                          double *im1, uint64_t batch, void *stream);
/* rows row_stride >= n doubles apart in each plane (packed through a workspace; see cfft_c64_fwd_strided) */
cfft_status cfft_f128_fwd_strided(const cfft_plan *plan, double *re0, double *re1, double *im0, double *im1,
                                  uint64_t row_stride, uint64_t batch, void *stream);
cfft_status cfft_f128_inv_strided(const cfft_plan *plan, double *re0, double *re1, double *im0, double *im1,
                                  uint64_t row_stride, uint64_t batch, void *stream);
/* fft128 counterpart of cfft_c64_host_multi */
cfft_status cfft_f128_host_multi(const cfft_plan *const *plans, int nplans, int op, double *re0, double *re1, double *im0,
                                 double *im1, uint64_t len, uint64_t batch);
/* host-memory versions; `len` = doubles per array, must equal batch * n */
cfft_status cfft_f128_fwd_host(const cfft_plan *plan, double *re0, double *re1, double *im0,
                               double *im1, uint64_t len, uint64_t batch);
cfft_status cfft_f128_inv_host(const cfft_plan *plan, double *re0, double *re1, double *im0,
                               double *im1, uint64_t len, uint64_t batch);
/* fwd immediately followed by inv on the device between one upload and one download */
cfft_status cfft_f128_fwd_inv_host(const cfft_plan *plan, double *re0, double *re1, double *im0,
                                   double *im1, uint64_t len, uint64_t batch);

/* ---- f128 operators around the transform ---------------------------- */

/* The scalar `f128` operators of src/fft128/f128_ops.rs applied element-wise to device arrays
 * (hi / lo planes), bit-exact with the reference's scalar functions:
 *   ADD add_f128_f128 :311-321, SUB sub_f128_f128 :360-370, MUL mul_f128_f128 :395-400,
 *   DIV div_f128_f128 :477-491, ADD_ESTIMATE :302-307, SUB_ESTIMATE :350-356, DIV_ESTIMATE :457-474.
 * out may alias an input.  Stream ordered on `device`. */
enum {
    CFFT_F128_ADD = 0, CFFT_F128_SUB, CFFT_F128_MUL, CFFT_F128_DIV,
    CFFT_F128_ADD_ESTIMATE, CFFT_F128_SUB_ESTIMATE, CFFT_F128_DIV_ESTIMATE,
    /* mixed-operand forms (the Add/Sub/Mul/Div<f64> impls, f128_ops.rs:48-230): the f64 operand's lo plane is ignored and may
     * be NULL.  add_f64_f128 / mul_f64_f128 are the f128_f64 forms with the operands swapped (:294-298, :388-391). */
    CFFT_F128_ADD_F128_F64,  /* :286-291 */
    CFFT_F128_SUB_F128_F64,  /* :331-336 */
    CFFT_F128_SUB_F64_F128,  /* :339-345 */
    CFFT_F128_MUL_F128_F64,  /* :380-385 */
    CFFT_F128_DIV_F128_F64,  /* :431-448 */
    CFFT_F128_DIV_F64_F128,  /* :451-454 */
    CFFT_F128_ADD_F64_F64,   /* :279-283 */
    CFFT_F128_SUB_F64_F64,   /* :324-328 */
    CFFT_F128_MUL_F64_F64,   /* :373-377 */
    CFFT_F128_DIV_F64_F64    /* :413-428 */
};
cfft_status cfft_f128_binary_op(int device, int op, const double *a_hi, const double *a_lo,
                                const double *b_hi, const double *b_lo, double *out_hi,
                                double *out_lo, uint64_t len, void *stream);

/* Unary operators of `f128` on device arrays, bit-exact with the reference's scalar functions:
 *   SQR sqr :404-409, ABS abs :506-511, NEG the Neg impl :232-238, IS_NAN is_nan :499-501 (out_hi = 1.0 / 0.0, out_lo = 0.0),
 *   SINCOSPI sincospi :514-575 + tables :578-618: out = sin(pi a), out2 = cos(pi a); the reference panics on inputs outside
 *   [-1, 1], here such elements become NaN and the call returns CFFT_EINVAL (this one operator synchronises the stream).
 * out2_* is only used by SINCOSPI.  `to_f64` (:494-496) is the hi plane itself.  out may alias the input. */
enum { CFFT_F128_SQR = 0, CFFT_F128_ABS, CFFT_F128_NEG, CFFT_F128_SINCOSPI, CFFT_F128_IS_NAN };
cfft_status cfft_f128_unary_op(int device, int op, const double *a_hi, const double *a_lo, double *out_hi, double *out_lo,
                               double *out2_hi, double *out2_lo, uint64_t len, void *stream);
/* PartialOrd / PartialEq of `f128` (f128_ops.rs:240-274) element-wise: out[i] = -1 Less, 0 Equal (== is exactly this case),
 * 1 Greater, 2 None (unordered: a NaN decided the comparison).  b_lo == NULL compares with the f64 values b_hi (:248-259, :267-274). */
cfft_status cfft_f128_compare(int device, const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo,
                              int8_t *out, uint64_t len, void *stream);

/* lhs <- (lhs * rhs) * factor, point-wise on planar double-double complex arrays: the step between
 * fwd and inv of a negacyclic product exactly as the reference's tests do it (scalar cplx_mul,
 * src/fft128/mod.rs:310-326, loop at :2033-2047; factor = 2 / N there). */
cfft_status cfft_f128_cplx_mul_scale(int device, double *l_re0, double *l_re1, double *l_im0,
                                     double *l_im1, const double *r_re0, const double *r_re1,
                                     const double *r_im0, const double *r_im1, double factor,
                                     uint64_t len, void *stream);

/* lhs <- inv( (fwd(lhs) * rhs) * factor ) for `batch` transforms in ONE call: the negacyclic polynomial product
 * exactly as the reference's own tests run it (src/fft128/mod.rs:2018-2053: Plan::fwd on the left operand, the loop
 * of scalar cplx_mul + scale at :2033-2047, Plan::inv), defined as -- and bit-identical to -- cfft_f128_fwd,
 * cfft_f128_cplx_mul_scale, cfft_f128_inv in sequence.  lhs: four planes of batch * n doubles, standard order in,
 * standard order out, in place.  rhs: Fourier-domain operand (already through cfft_f128_fwd, bit-reversed order), n
 * doubles per plane shared by every row when rhs_row_stride == 0, or batch * n per plane when rhs_row_stride == n.
 * n <= 4096 runs as one kernel (the tile never leaves shared memory between the last forward and the first inverse
 * pass); larger n as the three launches. */
cfft_status cfft_f128_fwd_mul_inv(const cfft_plan *plan, double *l_re0, double *l_re1, double *l_im0, double *l_im1,
                                  const double *r_re0, const double *r_re1, const double *r_im0,
                                  const double *r_im1, uint64_t rhs_row_stride, double factor, uint64_t batch,
                                  void *stream);

/* ---- c64 element-wise products in the Fourier domain ------------------------------------
 * "The only operations that are performed in the Fourier domain are elementwise" (README.md:10-17,
 * src/lib.rs:9-16): what a caller does between fwd and inv of a convolution / external product.  The
 * reference leaves it to the caller's `c64` arithmetic, so the semantics are those of num_complex's
 * `*` and `+` on Complex64 (the type src/lib.rs:84 re-exports): re = a.re*b.re - a.im*b.im,
 * im = a.re*b.im + a.im*b.re, every operation individually rounded (no FMA); bit-exact.  Element order is
 * irrelevant as long as both operands come from the same plan.  Device pointers, stream ordered. */
cfft_status cfft_c64_mul_assign(int device, void *lhs_dev, const void *rhs_dev, uint64_t len, void *stream);
/* acc[i] += a[i] * b[i]  (product as above, then a component-wise add) */
cfft_status cfft_c64_mul_add_assign(int device, void *acc_dev, const void *a_dev, const void *b_dev, uint64_t len,
                                    void *stream);

/* out[r] = inv( sum_{k < k_terms} fwd(a[r][k]) (.) b[r][k] ),  r < batch: a whole convolution / external-product
 * step (README.md:10-17 of the reference: forward transforms, element-wise products, one inverse transform) in
 * ONE call on device memory, defined as -- and bit-identical to -- the composition
 *     cfft_c64_fwd on every a[r][k];  acc = a[r][0] * b[r][0]  (cfft_c64_mul_assign);
 *     acc += a[r][k] * b[r][k] for k = 1, 2, ... in that order  (cfft_c64_mul_add_assign);  cfft_c64_inv(acc)
 * so the result is unnormalised exactly like fwd followed by inv (n x the convolution, src/unordered.rs:902-940).
 *   a    [batch][k_terms][n] c64, read only;
 *   b    Fourier-domain operand in THIS plan's order: [k_terms][n] shared by every row when b_row_stride == 0
 *        (e.g. one bootstrapping-key GGSW against a batch of ciphertexts), else row r starts at b + r * b_row_stride
 *        (in c64 elements, >= k_terms * n);
 *   out  [batch][n]; may be the same pointer as a only when k_terms == 1 (in place), never b.
 * Plans of the (Dif16, 256) family with 256 <= n <= 8192 (what Method::Measure selects on this library) run it as
 * one kernel that keeps the products and the running sum on the SM: (2 k_terms + 1) x 16 n bytes of HBM traffic
 * per row instead of (6 k_terms + 1) x 16 n for the separate calls (n = 8192: one launch per term with the partial
 * sum kept in out, 3 x 16 n per term); cfft_plan_has_fused_mul_kernel tells (also for fft128 plans and
 * cfft_f128_fwd_mul_inv).  Every other c64 plan (ordered plans included) runs the same arithmetic from its own
 * kernels through a stream-ordered workspace.  Stream ordered on the plan's device. */
cfft_status cfft_c64_fwd_mul_inv(const cfft_plan *plan, const void *a_dev, uint64_t k_terms, const void *b_dev,
                                 uint64_t b_row_stride, void *out_dev, uint64_t batch, void *stream);
/* The same with n_out outputs per row that share the forward transforms -- the GLWE external product of a caller such as
 * TFHE-rs, where every decomposed term feeds each of the glwe_dimension + 1 output polynomials against its own row of the
 * Fourier-domain key:    out[r][o] = inv( sum_{k < k_terms} fwd(a[r][k]) (.) b[r][k][o] ),   o < n_out.
 *   a [batch][k_terms][n];  b [k_terms][n_out][n] (b_row_stride == 0: shared by every row) or per row at
 *   b + r * b_row_stride (>= k_terms * n_out * n);  out [batch][n_out][n], overlapping neither a nor b.
 * Defined as -- and bit-identical to -- cfft_c64_fwd_mul_inv once per output.  n_out == 2 on (Dif16, 256) plans of
 * n = 512 / 1024 / 2048 (polynomial sizes 1024 .. 4096) runs as ONE kernel in which every forward transform runs once
 * (2 k + 2 transforms per row instead of 4 k + 2; cfft_plan_has_fused_mul2_kernel tells); everything else runs output by
 * output through the stream-ordered workspace. */
cfft_status cfft_c64_fwd_mul_inv_multi(const cfft_plan *plan, const void *a_dev, uint64_t k_terms, const void *b_dev,
                                       uint64_t b_row_stride, uint64_t n_out, void *out_dev, uint64_t batch, void *stream);
int cfft_plan_has_fused_mul2_kernel(const cfft_plan *plan);
/* acc[r] <- fwd(a[r]) (.) b[r]  (accumulate == 0)   or   acc[r] <- acc[r] + fwd(a[r]) (.) b[r]  (accumulate != 0), r < batch:
 * the forward transform and the element-wise multiply[-accumulate] into a FOURIER-DOMAIN accumulator (this plan's order)
 * in one call, without the inverse -- for loops that produce their terms one at a time or feed several accumulators
 * from one input (out_j += fwd(a_i) (.) B[i][j]), finished by cfft_c64_inv(acc).  Bit-identical to cfft_c64_fwd on a
 * copy of a followed by cfft_c64_mul_assign / cfft_c64_mul_add_assign.  Row r of a starts at a + r * a_row_stride
 * (c64 elements, a positive multiple of n: picks one term out of a [batch][k][n] array), row r of b at
 * b + r * b_row_stride (0 = shared), acc is [batch][n] and must not alias a or b.  One kernel on plans for which
 * cfft_plan_has_fused_mul_kernel answers 1 (a read once, acc read and written once), else copy + fwd + product. */
cfft_status cfft_c64_fwd_mul_add(const cfft_plan *plan, const void *a_dev, uint64_t a_row_stride, const void *b_dev,
                                 uint64_t b_row_stride, void *acc_dev, int accumulate, uint64_t batch, void *stream);
int cfft_plan_has_fused_mul_kernel(const cfft_plan *plan);

/* ---- integer polynomials <-> the Fourier domain --------------------------------
 * What a caller does on either side of the transforms of a negacyclic polynomial product modulo X^N + 1 (N = 2 n), fused
 * into the transform's first / last pass so that the conversions cost no HBM traffic of their own:
 *   in :  fold       z_j = coeff[j] + i coeff[j + n], j < n            (the fold of src/fft128/mod.rs:2006-2016)
 *         convert    signed 64-bit -> f64, round to nearest; CFFT_POLY_TORUS: the u64 torus element reinterpreted as i64, x 2^-64
 *         twist      z_j <- z_j * e^{+i pi j / N}, table entries (cos, sin) from sincospi64 (src/fft_simd.rs:237-296),
 *                    product with num_complex semantics (no FMA, src/lib.rs:84)
 *         then Plan::fwd
 *   out:  Plan::inv, then untwist + scale  t_j = z_j * (conj(twist_j) / n), then
 *         integer mode: coeff = f64::round(t) (half away from zero) as i64 (saturating, NaN -> 0)
 *         torus mode:   coeff = round((t - round(t)) * 2^64) modulo 2^64
 *         coeff[j] <- re, coeff[j + n] <- im; CFFT_POLY_ACCUMULATE adds to the existing coefficients modulo 2^64 instead.
 * The reference crate stops at the transform (README.md:10-17): these steps live in its caller, so the choices above are this
 * library's (documented here, restated on the CPU by the test oracle, and checked end to end against exact integer
 * schoolbook products); the transforms in between are the reference's, bit for bit.
 * Polynomials: 2 n int64 per row, 8-byte aligned; Fourier-domain buffers: n c64 per row, 16-byte aligned, in THIS plan's
 * order.  Plans of the (Dif16, 256) family with 256 <= n <= 8192 run each call as ONE kernel; every other c64 plan runs
 * stand-alone conversion kernels around its own transform (same bits).  Device pointers, stream ordered. */
enum { CFFT_POLY_INTEGER = 0, CFFT_POLY_TORUS = 1, CFFT_POLY_ACCUMULATE = 2 };
/* fourier[r] = fwd(twist(fold(poly[r]))), r < batch.  flags: CFFT_POLY_TORUS or 0. */
cfft_status cfft_c64_poly_fwd(const cfft_plan *plan, const int64_t *poly_dev, void *fourier_dev, uint64_t batch,
                              uint32_t flags, void *stream);
/* poly[r] (+)= round(untwist(inv(fourier[r]))); fourier is not modified. */
cfft_status cfft_c64_poly_inv(const cfft_plan *plan, const void *fourier_dev, int64_t *poly_dev, uint64_t batch,
                              uint32_t flags, void *stream);
/* out[r] (+)= round(untwist(inv( sum_{k < k_terms} fwd(twist(fold(a[r][k]))) (.) b[r][k] ))): a whole negacyclic
 * product / external-product step, integers in, integers out, one kernel for n <= 4096 (n = 8192: k_terms == 1).
 * a: [batch][k_terms][2n] int64; b: Fourier-domain operand as in cfft_c64_fwd_mul_inv (b_row_stride 0 = shared by every
 * row); out: [batch][2n] int64, may be a itself when k_terms == 1 and not accumulating.  Defined as -- and bit-identical
 * to -- cfft_c64_poly_fwd on every term, cfft_c64_mul_assign / cfft_c64_mul_add_assign in term order, cfft_c64_poly_inv. */
cfft_status cfft_c64_poly_mul(const cfft_plan *plan, const int64_t *a_dev, uint64_t k_terms, const void *b_dev,
                              uint64_t b_row_stride, int64_t *out_dev, uint64_t batch, uint32_t flags, void *stream);
/* Same with the polynomials in HOST memory (pinned or pageable) and b resident on the device (a bootstrapping / key-switching
 * key stays in the Fourier domain on the GPU): per row k_terms * 16 n bytes go up and 16 n come back, against
 * 2 * 16 n each way per TRANSFORM for the plain host entry points.  Synchronous. */
cfft_status cfft_c64_poly_mul_host(const cfft_plan *plan, const int64_t *a_host, uint64_t k_terms, const void *b_dev,
                                   uint64_t b_row_stride, int64_t *out_host, uint64_t batch, uint32_t flags);
/* 1 when the three calls above run as single fused kernels for this plan and term count */
int cfft_plan_has_fused_poly_kernel(const cfft_plan *plan, uint64_t k_terms);
/* the plan's twist tables: n entries e^{+i pi j / 2n}, then n entries conj / n (tests) */
cfft_status cfft_plan_copy_twist(const cfft_plan *plan, void *host_out, uint64_t bytes);

/* ---- diagnostics ------------------------------------------------------------------- */

const char *cfft_status_string(cfft_status st);
/* thread-local text of the last failure on this thread ("" if none) */
const char *cfft_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t cfft_launch_count(void);
/* "cfft_b200 <version> sm_100a" */
const char *cfft_version(void);
/* The opt-in persistent two-phase kernel (CFFT_B200_FAST_VARIANT=8 / CFFT_B200_ALLOW_PERSISTENT) waits on other CTAs with
 * a bounded spin; if a wait ever expires (producers descheduled by MPS, a debugger, preemption) the kernel gives up waiting
 * instead of hanging or trapping and counts it here: a non-zero value means such calls returned invalid data. */
cfft_status cfft_twopass_timeouts(int device, uint32_t *out);
/* Measured FP64 issue rate of `device` (thread-level FP64 instructions per second over the whole GPU): chains of
 * DFMA, of DADD, and a 1 : 7 DFMA : DADD mix like the fft128 butterfly's (94 instructions = 78 DADD + 12 DFMA + 4 DMUL,
 * src/fft128/mod.rs:310-346), 16 warps per SM, timed with CUDA events; plus the SM clock observed inside the kernel
 * (clock64 / globaltimer) and the SM count.  bench.py uses it as the measured peak of the fft128 roofline
 * ("measure it with a DFMA microbenchmark").  Any out pointer may be NULL. */
cfft_status cfft_probe_fp64_issue_rate(int device, double *dfma_per_s, double *dadd_per_s, double *mix_per_s,
                                       double *sm_mhz, int *sm_count);
/* copy of a plan's device twiddle table to host (tests: table parity with the reference's
 * init_wt / init_twiddles / init_negacyclic_twiddles).  which: c64 0 = fwd, 1 = inv table
 * (n + base_n c64 each; 2n for ordered); fft128 0..3 = re0, re1, im0, im1 (n doubles). */
cfft_status cfft_plan_copy_twiddles(const cfft_plan *plan, int which, void *host_out, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* CFFT_B200_H */

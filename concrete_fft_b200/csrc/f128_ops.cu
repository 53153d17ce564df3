// f128_ops.cu -- element-wise double-double ("f128") operators on the device, and the point-wise
// complex product used between fwd and inv in a negacyclic convolution.
//
// Bit-exact restatements of the reference's SCALAR operators (the Rust `f128` methods callers use
// around the transform): src/fft128/f128_ops.rs
//   add_f128_f128 :311-321   sub_f128_f128 :360-370   mul_f128_f128 :395-400   div_f128_f128 :477-491
//   add_estimate :302-307    sub_estimate :350-356    div_estimate_f128_f128 :457-474
//   helpers two_sum/two_diff/two_prod/quick_two_sum :6-40, add_f128_f64 :286-291, mul_f128_f64 :380-385
// and of the loop at src/fft128/mod.rs:2033-2047 (scalar cplx_mul :310-326, then a power-of-two scale).
// Pure streaming kernels: 48 B (binary op) / 96 B (complex product) of HBM traffic per element.
#include <cuda_runtime.h>

#include "plan.h"

namespace cfft {
namespace {

#define OPS_DEV __device__ __forceinline__
struct dd { double hi, lo; };

OPS_DEV dd quick_two_sum(double a, double b) { double s = __dadd_rn(a, b); return {s, __dsub_rn(b, __dsub_rn(s, a))}; }
OPS_DEV dd two_sum(double a, double b)
{
    double s = __dadd_rn(a, b), bb = __dsub_rn(s, a);
    return {s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
OPS_DEV dd two_diff(double a, double b)
{
    double s = __dsub_rn(a, b), bb = __dsub_rn(s, a);
    return {s, __dsub_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dadd_rn(b, bb))};
}
OPS_DEV dd two_prod(double a, double b) { double p = __dmul_rn(a, b); return {p, __fma_rn(a, b, -p)}; }

OPS_DEV dd add_est(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    return quick_two_sum(s.hi, __dadd_rn(s.lo, __dadd_rn(a.lo, b.lo)));
}
OPS_DEV dd sub_est(dd a, dd b)
{
    dd s = two_diff(a.hi, b.hi);
    return quick_two_sum(s.hi, __dsub_rn(__dadd_rn(s.lo, a.lo), b.lo));
}
OPS_DEV dd add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi), t = two_sum(a.lo, b.lo);
    s = quick_two_sum(s.hi, __dadd_rn(s.lo, t.hi));
    return quick_two_sum(s.hi, __dadd_rn(s.lo, t.lo));
}
OPS_DEV dd sub(dd a, dd b)
{
    dd s = two_diff(a.hi, b.hi), t = two_diff(a.lo, b.lo);
    s = quick_two_sum(s.hi, __dadd_rn(s.lo, t.hi));
    return quick_two_sum(s.hi, __dadd_rn(s.lo, t.lo));
}
OPS_DEV dd mul(dd a, dd b) // scalar form: p2 + (a0*b1 + a1*b0)
{
    dd p = two_prod(a.hi, b.hi);
    return quick_two_sum(p.hi, __dadd_rn(p.lo, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi))));
}
OPS_DEV dd mul_d(dd a, double b)
{
    dd p = two_prod(a.hi, b);
    return quick_two_sum(p.hi, __dadd_rn(p.lo, __dmul_rn(a.lo, b)));
}
OPS_DEV dd add_d(dd a, double b)
{
    dd s = two_sum(a.hi, b);
    return quick_two_sum(s.hi, __dadd_rn(s.lo, a.lo));
}
OPS_DEV dd div(dd a, dd b)
{
    const double q1 = __ddiv_rn(a.hi, b.hi);
    dd r = sub(a, mul_d(b, q1));
    const double q2 = __ddiv_rn(r.hi, b.hi);
    r = sub(r, mul_d(b, q2));
    const double q3 = __ddiv_rn(r.hi, b.hi);
    return add_d(quick_two_sum(q1, q2), q3);
}
OPS_DEV dd div_est(dd a, dd b)
{
    const double q1 = __ddiv_rn(a.hi, b.hi);
    const dd r = mul_d(b, q1);
    const dd s = two_diff(a.hi, r.hi);
    const double s2 = __dadd_rn(__dsub_rn(s.lo, r.lo), a.lo);
    const double q2 = __ddiv_rn(__dadd_rn(s.hi, s2), b.hi);
    return quick_two_sum(q1, q2);
}

// mixed-operand forms, f128_ops.rs:324-347 (sub), :413-455 (div)
OPS_DEV dd sub_f128_f64(dd a, double b)
{
    dd s = two_diff(a.hi, b);
    return quick_two_sum(s.hi, __dadd_rn(s.lo, a.lo));
}
OPS_DEV dd sub_f64_f128(double a, dd b)
{
    dd s = two_diff(a, b.hi);
    return quick_two_sum(s.hi, __dsub_rn(s.lo, b.lo));
}
OPS_DEV dd div_f64_f64(double a, double b)
{
    const double q1 = __ddiv_rn(a, b);
    const dd p = two_prod(q1, b);
    const dd s = two_diff(a, p.hi);
    const double e = __dsub_rn(s.lo, p.lo);
    return quick_two_sum(q1, __ddiv_rn(__dadd_rn(s.hi, e), b));
}
OPS_DEV dd div_f128_f64(dd a, double b)
{
    const double q1 = __ddiv_rn(a.hi, b);
    const dd p = two_prod(q1, b);
    const dd s = two_diff(a.hi, p.hi);
    const double e = __dsub_rn(__dadd_rn(s.lo, a.lo), p.lo);
    return quick_two_sum(q1, __ddiv_rn(__dadd_rn(s.hi, e), b));
}
OPS_DEV dd sqr(dd a) // :404-409
{
    dd p = two_prod(a.hi, a.hi);
    return quick_two_sum(p.hi, __dadd_rn(p.lo, __dmul_rn(2.0, __dmul_rn(a.hi, a.lo))));
}
OPS_DEV dd neg(dd a) { return {-a.hi, -a.lo}; }

// f128::sincospi, f128_ops.rs:514-618 (tables :578-618)
__constant__ double kSinTaylor[9][2] = {
    {-5.16771278004997, 2.2665622825789447e-16},     {2.5501640398773455, -7.931006345326556e-17},
    {-0.5992645293207921, 2.845026112698218e-17},    {0.08214588661112823, -3.847292805297656e-18},
    {-0.0073704309457143504, -3.328281165603432e-19}, {0.00046630280576761255, 1.0704561733683463e-20},
    {-2.1915353447830217e-5, 1.4648526682685598e-21}, {7.952054001475513e-7, 1.736540361519021e-23},
    {-2.2948428997269873e-8, -7.376346207041088e-26}};
__constant__ double kCosTaylor[9][2] = {
    {-4.934802200544679, -3.1326477543698557e-16},   {4.0587121264167685, -2.6602000824298645e-16},
    {-1.3352627688545895, 3.1815237892149862e-18},   {0.2353306303588932, -1.2583065576724427e-18},
    {-0.02580689139001406, 1.170191067939226e-18},   {0.0019295743094039231, -9.669517939986956e-20},
    {-0.0001046381049248457, -2.421206183964864e-21}, {4.303069587032947e-6, -2.864010082936791e-22},
    {-1.3878952462213771e-7, -7.479362090417238e-24}};
__constant__ double kSinK16[4][2] = {{0.19509032201612828, -7.991079068461731e-18}, {0.3826834323650898, -1.0050772696461588e-17},
                                     {0.5555702330196022, 4.709410940561677e-17},   {0.7071067811865476, -4.833646656726457e-17}};
__constant__ double kCosK16[4][2] = {{0.9807852804032304, 1.8546939997825006e-17}, {0.9238795325112867, 1.7645047084336677e-17},
                                     {0.8314696123025452, 1.4073856984728024e-18}, {0.7071067811865476, -4.833646656726457e-17}};

__device__ void sincospi(dd x, dd &s_out, dd &c_out)
{
    // approximately reduce modulo 1/2, then modulo 1/16 (:539-545)
    const double p = round(__dmul_rn(x.hi, 2.0));
    dd r = sub_f128_f64(x, __dmul_rn(p, 0.5));
    const double q = round(__dmul_rn(r.hi, 16.0));
    r = sub_f128_f64(r, __dmul_rn(q, 0.0625));
    // Taylor series in r (:514-532)
    dd sinc = {3.141592653589793, 1.2246467991473532e-16}, cosv = {1.0, 0.0}, pw = {1.0, 0.0};
    const dd sq = sqr(r);
    for (int i = 0; i < 9; i++) {
        pw = mul(pw, sq);
        sinc = add(sinc, mul(dd{kSinTaylor[i][0], kSinTaylor[i][1]}, pw));
        cosv = add(cosv, mul(dd{kCosTaylor[i][0], kCosTaylor[i][1]}, pw));
    }
    const dd sin_r = mul(sinc, r), cos_r = cosv;
    const int qi = int(q), pi = int(p);
    dd s, c;
    if (qi == 0) {
        s = sin_r;
        c = cos_r;
    } else {
        const int qa = (qi < 0 ? -qi : qi) - 1;
        const dd u = {kCosK16[qa][0], kCosK16[qa][1]}, v = {kSinK16[qa][0], kSinK16[qa][1]};
        if (qi > 0) {
            s = add(mul(u, sin_r), mul(v, cos_r));
            c = sub(mul(u, cos_r), mul(v, sin_r));
        } else {
            s = sub(mul(u, sin_r), mul(v, cos_r));
            c = add(mul(u, cos_r), mul(v, sin_r));
        }
    }
    if (pi == 0) { s_out = s; c_out = c; }
    else if (pi == 1) { s_out = c; c_out = neg(s); }
    else if (pi == -1) { s_out = neg(c); c_out = s; }
    else { s_out = neg(s); c_out = neg(c); }
}

// a_lo / b_lo may be null for the forms whose operand is an f64
template <int OP>
__global__ void f128_binary_kernel(const double *__restrict__ a_hi, const double *__restrict__ a_lo,
                                   const double *__restrict__ b_hi, const double *__restrict__ b_lo, double *out_hi,
                                   double *out_lo, uint64_t len)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += uint64_t(gridDim.x) * blockDim.x) {
        const dd a = {a_hi[i], a_lo ? a_lo[i] : 0.0}, b = {b_hi[i], b_lo ? b_lo[i] : 0.0};
        dd r;
        if (OP == 0) r = add(a, b);
        else if (OP == 1) r = sub(a, b);
        else if (OP == 2) r = mul(a, b);
        else if (OP == 3) r = div(a, b);
        else if (OP == 4) r = add_est(a, b);
        else if (OP == 5) r = sub_est(a, b);
        else if (OP == 6) r = div_est(a, b);
        else if (OP == 7) r = add_d(a, b.hi);
        else if (OP == 8) r = sub_f128_f64(a, b.hi);
        else if (OP == 9) r = sub_f64_f128(a.hi, b);
        else if (OP == 10) r = mul_d(a, b.hi);
        else if (OP == 11) r = div_f128_f64(a, b.hi);
        else if (OP == 12) r = div(dd{a.hi, 0.0}, b);
        else if (OP == 13) r = two_sum(a.hi, b.hi);
        else if (OP == 14) r = two_diff(a.hi, b.hi);
        else if (OP == 15) r = two_prod(a.hi, b.hi);
        else r = div_f64_f64(a.hi, b.hi);
        out_hi[i] = r.hi;
        out_lo[i] = r.lo;
    }
}

// 0 sqr, 1 abs, 2 neg, 3 sincospi (out = sin, out2 = cos; |x| > 1 -> NaN and *bad is raised), 4 is_nan (out_hi = 1 / 0)
template <int OP>
__global__ void f128_unary_kernel(const double *__restrict__ a_hi, const double *__restrict__ a_lo, double *out_hi, double *out_lo,
                                  double *out2_hi, double *out2_lo, uint64_t len, unsigned int *bad)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += uint64_t(gridDim.x) * blockDim.x) {
        const dd a = {a_hi[i], a_lo[i]};
        dd r = {0.0, 0.0};
        if (OP == 0) r = sqr(a);
        else if (OP == 1) r = a.hi < 0.0 ? neg(a) : a;
        else if (OP == 2) r = neg(a);
        else if (OP == 3) {
            dd c;
            // the reference panics outside [-1, 1] (f128 comparison with 1.0 / -1.0, f128_ops.rs:536-538)
            const bool gt = a.hi > 1.0 || (a.hi == 1.0 && a.lo > 0.0), lt = a.hi < -1.0 || (a.hi == -1.0 && a.lo < 0.0);
            if (gt || lt) {
                atomicAdd(bad, 1u);
                r = c = dd{__longlong_as_double(0x7FF8000000000000ll), __longlong_as_double(0x7FF8000000000000ll)};
            } else {
                sincospi(a, r, c);
            }
            out2_hi[i] = c.hi;
            out2_lo[i] = c.lo;
        } else {
            r.hi = (a.hi != a.hi || a.lo != a.lo) ? 1.0 : 0.0;
        }
        out_hi[i] = r.hi;
        out_lo[i] = r.lo;
    }
}

// PartialOrd, f128_ops.rs:240-274: -1 Less, 0 Equal, 1 Greater, 2 None; b_lo null: the right operand is an f64
__device__ __forceinline__ int cmp_f64(double x, double y) { return x < y ? -1 : (x > y ? 1 : (x == y ? 0 : 2)); }
__global__ void f128_compare_kernel(const double *__restrict__ a_hi, const double *__restrict__ a_lo, const double *__restrict__ b_hi,
                                    const double *__restrict__ b_lo, signed char *out, uint64_t len)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += uint64_t(gridDim.x) * blockDim.x) {
        const int first = cmp_f64(a_hi[i], b_hi[i]);
        out[i] = static_cast<signed char>(first == 0 ? cmp_f64(a_lo[i], b_lo ? b_lo[i] : 0.0) : first);
    }
}

// rhs_period != 0: the right operand holds rhs_period points shared by every row (index i mod rhs_period, a power of two)
__global__ void f128_cplx_mul_scale_kernel(double *l_re0, double *l_re1, double *l_im0, double *l_im1,
                                           const double *__restrict__ r_re0, const double *__restrict__ r_re1,
                                           const double *__restrict__ r_im0, const double *__restrict__ r_im1,
                                           double factor, uint64_t len, uint64_t rhs_period = 0)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t j = rhs_period ? (i & (rhs_period - 1)) : i;
        const dd ar = {l_re0[i], l_re1[i]}, ai = {l_im0[i], l_im1[i]};
        const dd br = {r_re0[j], r_re1[j]}, bi = {r_im0[j], r_im1[j]};
        const dd rr = mul(ar, br), ri = mul(ar, bi), ir = mul(ai, br), ii = mul(ai, bi);
        const dd pr = sub_est(rr, ii), pi = add_est(ir, ri);
        l_re0[i] = __dmul_rn(pr.hi, factor);
        l_re1[i] = __dmul_rn(pr.lo, factor);
        l_im0[i] = __dmul_rn(pi.hi, factor);
        l_im1[i] = __dmul_rn(pi.lo, factor);
    }
}

unsigned grid_for(uint64_t len)
{
    uint64_t blocks = (len + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    return unsigned(blocks ? blocks : 1);
}

} // namespace

cudaError_t launch_f128_binary(int op, const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo,
                               double *out_hi, double *out_lo, uint64_t len, cudaStream_t st)
{
    if (len == 0) return cudaSuccess;
    const unsigned g = grid_for(len);
    switch (op) {
#define CFFT_BIN(OPN) case OPN: f128_binary_kernel<OPN><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    CFFT_BIN(0) CFFT_BIN(1) CFFT_BIN(2) CFFT_BIN(3) CFFT_BIN(4) CFFT_BIN(5) CFFT_BIN(6) CFFT_BIN(7) CFFT_BIN(8)
    CFFT_BIN(9) CFFT_BIN(10) CFFT_BIN(11) CFFT_BIN(12) CFFT_BIN(13) CFFT_BIN(14) CFFT_BIN(15) CFFT_BIN(16)
#undef CFFT_BIN
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_f128_unary(int op, const double *a_hi, const double *a_lo, double *out_hi, double *out_lo, double *out2_hi,
                              double *out2_lo, uint64_t len, unsigned int *bad, cudaStream_t st)
{
    if (len == 0) return cudaSuccess;
    const unsigned g = grid_for(len);
    switch (op) {
    case 0: f128_unary_kernel<0><<<g, 256, 0, st>>>(a_hi, a_lo, out_hi, out_lo, out2_hi, out2_lo, len, bad); break;
    case 1: f128_unary_kernel<1><<<g, 256, 0, st>>>(a_hi, a_lo, out_hi, out_lo, out2_hi, out2_lo, len, bad); break;
    case 2: f128_unary_kernel<2><<<g, 256, 0, st>>>(a_hi, a_lo, out_hi, out_lo, out2_hi, out2_lo, len, bad); break;
    case 3: f128_unary_kernel<3><<<g, 256, 0, st>>>(a_hi, a_lo, out_hi, out_lo, out2_hi, out2_lo, len, bad); break;
    case 4: f128_unary_kernel<4><<<g, 256, 0, st>>>(a_hi, a_lo, out_hi, out_lo, out2_hi, out2_lo, len, bad); break;
    default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_f128_compare(const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo, signed char *out,
                                uint64_t len, cudaStream_t st)
{
    if (len == 0) return cudaSuccess;
    f128_compare_kernel<<<grid_for(len), 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out, len);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_f128_cplx_mul_scale(double *l_re0, double *l_re1, double *l_im0, double *l_im1, const double *r_re0,
                                       const double *r_re1, const double *r_im0, const double *r_im1, double factor,
                                       uint64_t len, cudaStream_t st)
{
    if (len == 0) return cudaSuccess;
    f128_cplx_mul_scale_kernel<<<grid_for(len), 256, 0, st>>>(l_re0, l_re1, l_im0, l_im1, r_re0, r_re1, r_im0, r_im1,
                                                               factor, len);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_f128_cplx_mul_scale_rows(double *l_re0, double *l_re1, double *l_im0, double *l_im1, const double *r_re0,
                                            const double *r_re1, const double *r_im0, const double *r_im1, uint64_t rhs_period,
                                            double factor, uint64_t len, cudaStream_t st)
{
    if (len == 0) return cudaSuccess;
    f128_cplx_mul_scale_kernel<<<grid_for(len), 256, 0, st>>>(l_re0, l_re1, l_im0, l_im1, r_re0, r_re1, r_im0, r_im1,
                                                               factor, len, rhs_period);
    count_launch();
    return cudaGetLastError();
}

} // namespace cfft

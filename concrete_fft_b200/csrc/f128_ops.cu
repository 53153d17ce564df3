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

template <int OP>
__global__ void f128_binary_kernel(const double *__restrict__ a_hi, const double *__restrict__ a_lo,
                                   const double *__restrict__ b_hi, const double *__restrict__ b_lo, double *out_hi,
                                   double *out_lo, uint64_t len)
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += uint64_t(gridDim.x) * blockDim.x) {
        const dd a = {a_hi[i], a_lo[i]}, b = {b_hi[i], b_lo[i]};
        dd r;
        if (OP == 0) r = add(a, b);
        else if (OP == 1) r = sub(a, b);
        else if (OP == 2) r = mul(a, b);
        else if (OP == 3) r = div(a, b);
        else if (OP == 4) r = add_est(a, b);
        else if (OP == 5) r = sub_est(a, b);
        else r = div_est(a, b);
        out_hi[i] = r.hi;
        out_lo[i] = r.lo;
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
    case 0: f128_binary_kernel<0><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    case 1: f128_binary_kernel<1><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    case 2: f128_binary_kernel<2><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    case 3: f128_binary_kernel<3><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    case 4: f128_binary_kernel<4><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    case 5: f128_binary_kernel<5><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    case 6: f128_binary_kernel<6><<<g, 256, 0, st>>>(a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len); break;
    default: return cudaErrorInvalidValue;
    }
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

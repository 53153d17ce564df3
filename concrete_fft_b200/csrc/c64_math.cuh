// c64_math.cuh -- device complex-f64 primitives and radix-2/4/8/16 butterflies.
//
// Every operation is an explicitly rounded IEEE double op (__dadd_rn / __dsub_rn / __dmul_rn /
// __fma_rn), so nvcc can neither contract nor re-associate anything: the butterflies below
// produce the same bits as the reference's (scalar == AVX2 == AVX-512) butterflies:
//   complex multiply   src/fft_simd.rs:220-233  (re = fma(a,x,-(b*y)), im = fma(a,y,b*x))
//   mul_j, e^{+-i pi/4}, e^{+-i pi/8} family  src/fft_simd.rs:106-160
//   radix-2  src/dif2.rs:106-113      radix-4  src/dif4.rs:195-214
//   radix-8  src/dif8.rs:310-351 (== src/unordered.rs:98-219)
//   radix-16 src/dif16.rs:649-772
#pragma once
#include <cuda_runtime.h>

namespace cfft {

typedef double2 c64; // .x = re, .y = im  (src/lib.rs:84)

#define CFFT_DEV __device__ __forceinline__

CFFT_DEV c64 mk(double re, double im) { return make_double2(re, im); }
CFFT_DEV c64 cadd(c64 a, c64 b) { return mk(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
CFFT_DEV c64 csub(c64 a, c64 b) { return mk(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }

// w * z with the reference's rounding
CFFT_DEV c64 cmul(c64 w, c64 z)
{
    return mk(__fma_rn(w.x, z.x, -__dmul_rn(w.y, z.y)), __fma_rn(w.x, z.y, __dmul_rn(w.y, z.x)));
}

template <bool FWD> CFFT_DEV c64 mulj(c64 z) { return FWD ? mk(-z.y, z.x) : mk(z.y, -z.x); }

// e^{-+ i pi/4}-type rotations: (1/sqrt2) * (z + (+-j) z)
template <bool FWD> CFFT_DEV c64 mul_e8(c64 z)
{
    const double r = 0.7071067811865476;
    c64 t = cadd(z, mulj<FWD>(z));
    return mk(__dmul_rn(r, t.x), __dmul_rn(r, t.y));
}
template <bool FWD> CFFT_DEV c64 mul_ne8(c64 z) { return mul_e8<!FWD>(z); }

#define CFFT_H1X 0.9238795325112867
#define CFFT_H1Y (-0.38268343236508984)
template <bool FWD> CFFT_DEV c64 mul_e16(c64 z) { return cmul(mk(CFFT_H1X, FWD ? CFFT_H1Y : -CFFT_H1Y), z); }
template <bool FWD> CFFT_DEV c64 mul_e17(c64 z) { return cmul(mk(-CFFT_H1Y, FWD ? -CFFT_H1X : CFFT_H1X), z); }
template <bool FWD> CFFT_DEV c64 mul_ne16(c64 z) { return mul_e16<!FWD>(z); }
template <bool FWD> CFFT_DEV c64 mul_ne17(c64 z) { return mul_e17<!FWD>(z); }

CFFT_DEV void bf2(c64 *v)
{
    c64 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}

template <bool FWD> CFFT_DEV void bf4(c64 *v)
{
    c64 apc = cadd(v[0], v[2]), amc = csub(v[0], v[2]);
    c64 bpd = cadd(v[1], v[3]), jbmd = mulj<FWD>(csub(v[1], v[3]));
    v[0] = cadd(apc, bpd);
    v[1] = csub(amc, jbmd);
    v[2] = csub(apc, bpd);
    v[3] = cadd(amc, jbmd);
}

template <bool FWD> CFFT_DEV void bf8(c64 *v)
{
    c64 a0 = cadd(v[0], v[4]), s0 = csub(v[0], v[4]);
    c64 a1 = cadd(v[1], v[5]), s1 = csub(v[1], v[5]);
    c64 a2 = cadd(v[2], v[6]), js2 = mulj<FWD>(csub(v[2], v[6]));
    c64 a3 = cadd(v[3], v[7]), js3 = mulj<FWD>(csub(v[3], v[7]));

    c64 a02p = cadd(a0, a2), s02m = csub(s0, js2), a02m = csub(a0, a2), s02p = cadd(s0, js2);
    c64 a13p = cadd(a1, a3);
    c64 w8 = mul_ne8<FWD>(csub(s1, js3));
    c64 j13 = mulj<FWD>(csub(a1, a3));
    c64 v8 = mul_e8<FWD>(cadd(s1, js3));

    v[0] = cadd(a02p, a13p);
    v[1] = cadd(s02m, w8);
    v[2] = csub(a02m, j13);
    v[3] = csub(s02p, v8);
    v[4] = csub(a02p, a13p);
    v[5] = csub(s02m, w8);
    v[6] = cadd(a02m, j13);
    v[7] = cadd(s02p, v8);
}

// half of the radix-16 butterfly: combines the (e, o) = (i, i+2) column pair, see bf16
template <bool FWD>
CFFT_DEV void bf16_half(c64 ape, c64 sme, c64 ame, c64 spe, c64 apo, c64 smo, c64 amo, c64 spo, c64 *t)
{
    c64 w8 = mul_ne8<FWD>(smo), j_ = mulj<FWD>(amo), v8 = mul_e8<FWD>(spo);
    t[0] = cadd(ape, apo);
    t[1] = cadd(sme, w8);
    t[2] = csub(ame, j_);
    t[3] = csub(spe, v8);
    t[4] = csub(ape, apo);
    t[5] = csub(sme, w8);
    t[6] = cadd(ame, j_);
    t[7] = cadd(spe, v8);
}

template <bool FWD> CFFT_DEV void bf16(c64 *v)
{
    c64 ap[4], sm[4], am[4], sp[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        c64 a_lo = cadd(v[i], v[i + 8]), s_lo = csub(v[i], v[i + 8]);
        c64 a_hi = cadd(v[i + 4], v[i + 12]), js = mulj<FWD>(csub(v[i + 4], v[i + 12]));
        ap[i] = cadd(a_lo, a_hi);
        sm[i] = csub(s_lo, js);
        am[i] = csub(a_lo, a_hi);
        sp[i] = cadd(s_lo, js);
    }
    c64 E[8], O[8];
    bf16_half<FWD>(ap[0], sm[0], am[0], sp[0], ap[2], sm[2], am[2], sp[2], E);
    bf16_half<FWD>(ap[1], sm[1], am[1], sp[1], ap[3], sm[3], am[3], sp[3], O);

    c64 u1 = mul_e16<FWD>(O[1]);
    c64 u2 = mul_ne8<FWD>(O[2]);
    c64 u3 = mul_e17<FWD>(O[3]);
    c64 u4 = mulj<FWD>(O[4]);
    c64 u5 = mul_ne17<FWD>(O[5]);
    c64 u6 = mul_e8<FWD>(O[6]);
    c64 u7 = mul_ne16<FWD>(O[7]);

    v[0] = cadd(E[0], O[0]);  v[8] = csub(E[0], O[0]);
    v[1] = cadd(E[1], u1);    v[9] = csub(E[1], u1);
    v[2] = cadd(E[2], u2);    v[10] = csub(E[2], u2);
    v[3] = cadd(E[3], u3);    v[11] = csub(E[3], u3);
    v[4] = csub(E[4], u4);    v[12] = cadd(E[4], u4);
    v[5] = csub(E[5], u5);    v[13] = cadd(E[5], u5);
    v[6] = csub(E[6], u6);    v[14] = cadd(E[6], u6);
    v[7] = csub(E[7], u7);    v[15] = cadd(E[7], u7);
}

template <int R, bool FWD> CFFT_DEV void bfR(c64 *v)
{
    if (R == 2) bf2(v);
    else if (R == 4) bf4<FWD>(v);
    else if (R == 8) bf8<FWD>(v);
    else bf16<FWD>(v);
}

} // namespace cfft

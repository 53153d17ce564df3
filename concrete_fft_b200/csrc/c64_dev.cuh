// c64_dev.cuh -- device helpers shared by the c64 register kernels (c64_fast.cu, c64_column.cu,
// c64_ord16.cu): streaming 128-bit HBM accesses, twiddle loads, and the 256-point Dif16 base FFT
// of one half-warp (src/dif16.rs:449-827).
#pragma once
#include <cstdint>

#include "c64_math.cuh"

namespace cfft {
namespace dev {

__device__ __forceinline__ c64 ld_stream(const c64 *p)
{
    c64 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(c64 *p, c64 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ c64 ld_tw(const c64 *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

// num_complex `*` (src/lib.rs:84 re-exports the type): four products, one subtraction, one addition, no FMA --
// the arithmetic of everything a caller does between / around the transforms (README.md:10-17)
__device__ __forceinline__ c64 cmul_nc(c64 x, c64 y)
{
    return mk(__dsub_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y)), __dadd_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x)));
}

// ---- how a register kernel sees one row in global memory ------------------------------------------------
// PIN / POUT = false: n c64, 128-bit streaming accesses (the transform API).
// PIN  = true: the row is a POLYNOMIAL of 2n signed 64-bit coefficients (cfft_c64_poly_*): element
//   pos is the fold (src/fft128/mod.rs:2006-2016) re = coeff[pos], im = coeff[pos + n], converted to f64 (torus mode:
//   x 2^-64), times the negacyclic twist e^{+i pi pos / 2n} -- all fused into the load.
// POUT = true: element pos is multiplied by the untwist conj(twist) / n, rounded half away from zero like f64::round
//   (torus mode: the fractional part x 2^64, modulo 2^64) and stored -- or added, modulo 2^64 -- as coeff[pos] and
//   coeff[pos + n].
enum { POLY_TORUS = 1, POLY_ACCUMULATE = 2 };

__device__ __forceinline__ unsigned long long poly_to_integer(double x)
{
    if (x != x) return 0;                                             // NaN -> 0 like Rust's `as i64`
    return static_cast<unsigned long long>(__double2ll_rn(round(x))); // the conversion saturates like Rust's `as i64`
}
__device__ __forceinline__ unsigned long long poly_to_torus(double x)
{
    double f = __dsub_rn(x, round(x));
    f = round(__dmul_rn(f, 18446744073709551616.0));
    if (f != f) return 0; // NaN / infinite input
    if (f >= 9223372036854775808.0) return 1ull << 63; // exactly 2^63: the same torus element as -2^63
    return static_cast<unsigned long long>(__double2ll_rn(f));
}

template <bool PIN, bool POUT> struct RowIo {
    const c64 *in;
    c64 *out;
    const long long *pin;
    long long *pout;
    const c64 *twist; // [0, n): twist, [n, 2n): untwist
    uint32_t n;
    uint32_t flags;
    __device__ __forceinline__ c64 ld(int pos) const
    {
        if (!PIN) return ld_stream(in + pos);
        const double scale = (flags & POLY_TORUS) ? 5.421010862427522e-20 /* 2^-64 */ : 1.0;
        const c64 z = mk(__dmul_rn(__ll2double_rn(__ldg(pin + pos)), scale), __dmul_rn(__ll2double_rn(__ldg(pin + n + pos)), scale));
        return cmul_nc(z, ld_tw(twist + pos));
    }
    __device__ __forceinline__ void st(int pos, c64 v) const
    {
        if (!POUT) {
            st_stream(out + pos, v);
            return;
        }
        const c64 t = cmul_nc(v, ld_tw(twist + n + pos));
        unsigned long long re, im;
        if (flags & POLY_TORUS) {
            re = poly_to_torus(t.x);
            im = poly_to_torus(t.y);
        } else {
            re = poly_to_integer(t.x);
            im = poly_to_integer(t.y);
        }
        if (flags & POLY_ACCUMULATE) {
            re += static_cast<unsigned long long>(pout[pos]);
            im += static_cast<unsigned long long>(pout[n + pos]);
        }
        pout[pos] = static_cast<long long>(re);
        pout[n + pos] = static_cast<long long>(im);
    }
};

// the batch a kernel works on: row r of the c64 side starts r * row_in / r * row_out elements in, of the polynomial
// side r * prow_in / r * prow_out coefficients in
template <bool PIN, bool POUT> struct BatchIo {
    const c64 *in;
    c64 *out;
    const long long *pin;
    long long *pout;
    const c64 *twist;
    uint64_t row_in, row_out, prow_in, prow_out;
    uint32_t n, flags;
    uint32_t ahead; // plain c64 input only: rows ahead whose lines a kernel may request into L2 (0 = never)
    __device__ __forceinline__ RowIo<PIN, POUT> row(uint64_t r) const
    {
        RowIo<PIN, POUT> io;
        io.in = PIN ? nullptr : in + r * row_in;
        io.out = POUT ? nullptr : out + r * row_out;
        io.pin = PIN ? pin + r * prow_in : nullptr;
        io.pout = POUT ? pout + r * prow_out : nullptr;
        io.twist = twist;
        io.n = n;
        io.flags = flags;
        return io;
    }
};
// host side: the plain c64 batch (rows row_in / row_out elements apart)
inline BatchIo<false, false> plain_batch(const c64 *in, c64 *out, uint64_t row_in, uint64_t row_out)
{
    BatchIo<false, false> b;
    b.in = in;
    b.out = out;
    b.pin = nullptr;
    b.pout = nullptr;
    b.twist = nullptr;
    b.row_in = row_in;
    b.row_out = row_out;
    b.prow_in = b.prow_out = 0;
    b.n = 0;
    b.flags = 0;
    b.ahead = 0;
    return b;
}
typedef RowIo<false, false> PlainRow;
__device__ __forceinline__ PlainRow plain_row(const c64 *in, c64 *out)
{
    PlainRow io;
    io.in = in;
    io.out = out;
    io.pin = nullptr;
    io.pout = nullptr;
    io.twist = nullptr;
    io.n = 0;
    io.flags = 0;
    return io;
}

template <int R> __device__ __forceinline__ constexpr int brev_c(int k)
{
    return R == 2 ? k : (R == 4 ? ((k & 1) << 1) | (k >> 1) : ((k & 1) << 2) | (k & 2) | (k >> 2));
}

// 256-point base FFT (Dif16: radix-16 s=1 with twiddles, then radix-16 end) of the half-warp
// that owns block `blk`; thread lane16 = p (first pass) = j (second pass).
// FWD selects the butterfly direction only; the table passed in is the direction's table.
// `sw_in` / `sw_out` (0..7) XOR the natural-order shared-memory positions read / written; the standard-order
// kernels use it so that the transposing pass that follows / precedes is bank-conflict free.
// `io` + `goff`: the row accessor and the block's first element inside the row, used when G_IN / G_OUT.
template <bool FWD, bool G_IN, bool G_OUT, class Io>
__device__ __forceinline__ void base256_io(const Io &io, int goff, const c64 *__restrict__ src, c64 *__restrict__ sm_blk, c64 *__restrict__ dst,
                                           const c64 *__restrict__ tw_planar, int lane16, c64 (&v)[16], int sw_in = 0, int sw_out = 0)
{
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16); // the 16 lanes that own this block
    // pass 1: x[p + 16k] -> y[16p + k] = w[p + 16k] * DFT16(x)_k       src/dif16.rs:449-623
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = G_IN ? io.ld(goff + lane16 + 16 * k) : src[(lane16 + 16 * k) ^ sw_in];
    bf16<FWD>(v);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw_planar + lane16 + 16 * k), v[k]);
    __syncwarp(hmask); // the half-warp has finished reading its block
#pragma unroll
    for (int k = 0; k < 16; k++) sm_blk[16 * lane16 + (k ^ lane16)] = v[k]; // XOR swizzle: conflict-free
    __syncwarp(hmask);
    // pass 2: terminal radix-16 on y[j + 16k]                           src/dif16.rs:649-827
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = sm_blk[16 * k + (lane16 ^ k)];
    bf16<FWD>(v);
    if (G_OUT) {
#pragma unroll
        for (int k = 0; k < 16; k++) io.st(goff + lane16 + 16 * k, v[k]);
    } else {
        __syncwarp(hmask); // swizzled data consumed by the whole half-warp before natural-order overwrite
#pragma unroll
        for (int k = 0; k < 16; k++) dst[(lane16 + 16 * k) ^ sw_out] = v[k];
    }
}

// pointer form: `src` / `dst` are global pointers to the block when G_IN / G_OUT, shared-memory pointers otherwise
template <bool FWD, bool G_IN, bool G_OUT>
__device__ __forceinline__ void base256(const c64 *__restrict__ src, c64 *__restrict__ sm_blk, c64 *__restrict__ dst,
                                        const c64 *__restrict__ tw_planar, int lane16, c64 (&v)[16], int sw_in = 0, int sw_out = 0)
{
    base256_io<FWD, G_IN, G_OUT>(plain_row(G_IN ? src : nullptr, G_OUT ? dst : nullptr), 0, src, sm_blk, dst, tw_planar, lane16, v, sw_in, sw_out);
}

} // namespace dev
} // namespace cfft

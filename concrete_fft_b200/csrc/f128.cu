// f128.cu -- fft128: negacyclic radix-2 transform on double-double ("f128") complex data,
// four planar f64 arrays (re hi, re lo, im hi, im lo), in place, bit-reversed Fourier order.
//
// Reference: src/fft128/mod.rs:352-402 (forward), :1105-1155 (inverse), complex multiply
// :310-346, double-double ops src/fft128/f128_ops.rs:6-40, 302-307, 350-356 and the FMA
// multiply :837-841 (the form the reference's AVX2 / AVX-512 paths execute on x86).
// Every double op is an explicitly rounded intrinsic, in the reference's order, so the output
// is bit-identical to the reference's AVX path.
//
// Work split: a thread owns G = 2^S elements (S <= 3 consecutive stages) at stride u and runs
// those stages in registers: 94 FP64 instructions per butterfly, no memory traffic in between.
// Passes exchange data through shared memory (one tile = <= 4096 elements = 128 KiB); the
// first pass reads HBM and the last writes HBM, so a transform moves 2 x 32 B x n and is bound
// by the FP64 pipe.  Stages whose span exceeds a tile (n > 4096) run as separate HBM passes.
#include <cuda_runtime.h>

#include <cstdlib>

#include "plan.h"

namespace cfft {
namespace {

#define F128_DEV __device__ __forceinline__

struct dd { double hi, lo; };

F128_DEV dd quick_two_sum(double a, double b)
{
    double s = __dadd_rn(a, b);
    return {s, __dsub_rn(b, __dsub_rn(s, a))};
}
F128_DEV dd two_sum(double a, double b)
{
    double s = __dadd_rn(a, b);
    double bb = __dsub_rn(s, a);
    return {s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
F128_DEV dd two_diff(double a, double b)
{
    double s = __dsub_rn(a, b);
    double bb = __dsub_rn(s, a);
    return {s, __dsub_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dadd_rn(b, bb))};
}
// add_estimate_f128_f128, f128_ops.rs:302-307
F128_DEV dd dd_add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    double e = __dadd_rn(s.lo, __dadd_rn(a.lo, b.lo));
    return quick_two_sum(s.hi, e);
}
// sub_estimate_f128_f128, f128_ops.rs:350-356
F128_DEV dd dd_sub(dd a, dd b)
{
    dd s = two_diff(a.hi, b.hi);
    double e = __dadd_rn(s.lo, a.lo);
    e = __dsub_rn(e, b.lo);
    return quick_two_sum(s.hi, e);
}
// mul_f128x4, f128_ops.rs:837-841
F128_DEV dd dd_mul(dd a, dd b)
{
    double p = __dmul_rn(a.hi, b.hi);
    double e = __fma_rn(a.hi, b.hi, -p);
    e = __fma_rn(a.hi, b.lo, __fma_rn(a.lo, b.hi, e));
    return quick_two_sum(p, e);
}

struct ddc { dd re, im; };

// forward butterfly: z1w = z1 * w; (z0 + z1w, z0 - z1w)            mod.rs:386-395, 310-326
F128_DEV void bfly_fwd(ddc &z0, ddc &z1, ddc w)
{
    dd rr = dd_mul(z1.re, w.re), ri = dd_mul(z1.re, w.im);
    dd ir = dd_mul(z1.im, w.re), ii = dd_mul(z1.im, w.im);
    dd zr = dd_sub(rr, ii), zi = dd_add(ir, ri);
    ddc a = {dd_add(z0.re, zr), dd_add(z0.im, zi)};
    ddc b = {dd_sub(z0.re, zr), dd_sub(z0.im, zi)};
    z0 = a;
    z1 = b;
}
// inverse butterfly: (z0 + z1, (z0 - z1) * conj(w))                mod.rs:1139-1148, 330-346
F128_DEV void bfly_inv(ddc &z0, ddc &z1, ddc w)
{
    dd dr = dd_sub(z0.re, z1.re), di = dd_sub(z0.im, z1.im);
    ddc a = {dd_add(z0.re, z1.re), dd_add(z0.im, z1.im)};
    dd rr = dd_mul(dr, w.re), ri = dd_mul(dr, w.im);
    dd ir = dd_mul(di, w.re), ii = dd_mul(di, w.im);
    z0 = a;
    z1 = {dd_add(rr, ii), dd_sub(ir, ri)};
}

struct Planes { double *p[4]; };

// twiddle k as one 32-byte record {re hi, re lo, im hi, im lo}: two 128-bit loads
struct __align__(32) Tw4 { double re_hi, re_lo, im_hi, im_lo; };

F128_DEV ddc load_tw(const Tw4 *__restrict__ tw, uint32_t i)
{
    const double2 a = __ldg(reinterpret_cast<const double2 *>(tw + i));
    const double2 b = __ldg(reinterpret_cast<const double2 *>(tw + i) + 1);
    return {{a.x, a.y}, {b.x, b.y}};
}

// Stages d0 .. d0+S-1 (t_d = n >> (d+1), m_d = 1 << d) on the G = 2^S elements
// base + e*u, u = n >> (d0+S).  `row_pos` = position of element 0 inside its transform.
template <int S, bool FWD>
F128_DEV void run_group(ddc (&z)[1 << S], const Tw4 *__restrict__ tw, uint32_t row_pos, uint32_t logn, int d0)
{
    constexpr int G = 1 << S;
    const uint32_t B = row_pos >> (logn - d0); // block index at stage d0
    if (FWD) {
#pragma unroll
        for (int k = 0; k < S; k++) {
            const int h = G >> (k + 1);
            const uint32_t m = 1u << (d0 + k);
#pragma unroll
            for (int blk = 0; blk < (1 << k); blk++) {
                const ddc w = load_tw(tw, m + (B << k) + blk);
#pragma unroll
                for (int j = 0; j < h; j++) bfly_fwd(z[blk * 2 * h + j], z[blk * 2 * h + j + h], w);
            }
        }
    } else {
#pragma unroll
        for (int k = S - 1; k >= 0; k--) {
            const int h = G >> (k + 1);
            const uint32_t m = 1u << (d0 + k);
#pragma unroll
            for (int blk = 0; blk < (1 << k); blk++) {
                const ddc w = load_tw(tw, m + (B << k) + blk);
#pragma unroll
                for (int j = 0; j < h; j++) bfly_inv(z[blk * 2 * h + j], z[blk * 2 * h + j + h], w);
            }
        }
    }
}

constexpr int kThreads = 256;

// Shared-memory tile: two arrays of double2 -- re = {hi, lo} and im = {hi, lo} per element -- so an
// element moves with two 128-bit accesses.  Index swizzle i ^ ((i >> 3) & 7): for every pass
// stride u = 2^a the eight lanes of a 128-bit phase land in eight different 16-byte banks.
F128_DEV uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u); }

// one pass over a tile; G_IN / G_OUT: that side is the HBM tile (four planar arrays), else smem.
// Groups are dealt to warps in contiguous runs (warp w owns groups [w gpw, (w+1) gpw) of the FULL tile),
// so a pass whose groups span G u <= full_tile / warps elements stays inside the warp's own contiguous
// block of the tile: consecutive such passes need no block barrier (see pass_is_warp_local).
template <int S, bool FWD, bool G_IN, bool G_OUT, int NT>
F128_DEV void tile_pass(const Planes &g, double2 *__restrict__ sre, double2 *__restrict__ sim, uint32_t tile,
                        uint32_t full_tile, uint32_t tile_row_off, uint32_t n, uint32_t logn, int d0,
                        const Tw4 *__restrict__ tw, bool vec)
{
    constexpr int G = 1 << S;
    const uint32_t u = n >> (d0 + S);
    // groups per warp, rounded up to whole 32-lane rounds (exact for power-of-two tiles >= 256 per warp)
    const uint32_t gpw = ((full_tile / G + (NT / 32) - 1) / (NT / 32) + 31u) & ~31u;
    const uint32_t g0 = (threadIdx.x >> 5) * gpw + (threadIdx.x & 31);
    for (uint32_t grp = g0; grp < g0 + gpw; grp += 32) {
        if (grp >= tile / G) break;
        const uint32_t hi = grp / u, lo = grp - hi * u;
        const uint32_t base = hi * (G * u) + lo;
        ddc z[G];
        if (G_IN) {
            if (vec && u == 1 && G >= 2) { // G consecutive doubles per plane: 128-bit loads
                double buf[4][G];
#pragma unroll
                for (int pl = 0; pl < 4; pl++)
#pragma unroll
                    for (int e = 0; e < G; e += 2) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(g.p[pl] + base + e);
                        buf[pl][e] = t2.x;
                        buf[pl][e + 1] = t2.y;
                    }
#pragma unroll
                for (int e = 0; e < G; e++) z[e] = {{buf[0][e], buf[1][e]}, {buf[2][e], buf[3][e]}};
            } else {
#pragma unroll
                for (int e = 0; e < G; e++) {
                    const uint32_t i = base + e * u;
                    z[e] = {{g.p[0][i], g.p[1][i]}, {g.p[2][i], g.p[3][i]}};
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < G; e++) {
                const uint32_t i = swz(base + e * u);
                const double2 re2 = sre[i], im2 = sim[i];
                z[e].re.hi = re2.x; z[e].re.lo = re2.y;
                z[e].im.hi = im2.x; z[e].im.lo = im2.y;
            }
        }
        run_group<S, FWD>(z, tw, (tile_row_off + base) & (n - 1), logn, d0);
        if (G_OUT) {
            if (vec && u == 1 && G >= 2) {
#pragma unroll
                for (int e = 0; e < G; e += 2) {
                    *reinterpret_cast<double2 *>(g.p[0] + base + e) = make_double2(z[e].re.hi, z[e + 1].re.hi);
                    *reinterpret_cast<double2 *>(g.p[1] + base + e) = make_double2(z[e].re.lo, z[e + 1].re.lo);
                    *reinterpret_cast<double2 *>(g.p[2] + base + e) = make_double2(z[e].im.hi, z[e + 1].im.hi);
                    *reinterpret_cast<double2 *>(g.p[3] + base + e) = make_double2(z[e].im.lo, z[e + 1].im.lo);
                }
            } else {
#pragma unroll
                for (int e = 0; e < G; e++) {
                    const uint32_t i = base + e * u;
                    g.p[0][i] = z[e].re.hi;
                    g.p[1][i] = z[e].re.lo;
                    g.p[2][i] = z[e].im.hi;
                    g.p[3][i] = z[e].im.lo;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < G; e++) {
                const uint32_t i = swz(base + e * u);
                sre[i] = make_double2(z[e].re.hi, z[e].re.lo);
                sim[i] = make_double2(z[e].im.hi, z[e].im.lo);
            }
        }
    }
}

constexpr int kMaxPasses = 6; // tile <= 4096 elements = 12 stages = 4 passes of 3 (or 6 of 2)
struct PassList {
    int count;
    int d0[kMaxPasses];
    int s[kMaxPasses];
};

template <bool FWD, bool G_IN, bool G_OUT, int NT, int SMAX = 3>
F128_DEV void dispatch_pass(int s, const Planes &g, double2 *sre, double2 *sim, uint32_t tile, uint32_t full_tile,
                            uint32_t row_off, uint32_t n, uint32_t logn, int d0, const Tw4 *tw, bool vec)
{
    if (SMAX >= 3 && s == 3) tile_pass<3, FWD, G_IN, G_OUT, NT>(g, sre, sim, tile, full_tile, row_off, n, logn, d0, tw, vec);
    else if (s == 2) tile_pass<2, FWD, G_IN, G_OUT, NT>(g, sre, sim, tile, full_tile, row_off, n, logn, d0, tw, vec);
    else tile_pass<1, FWD, G_IN, G_OUT, NT>(g, sre, sim, tile, full_tile, row_off, n, logn, d0, tw, vec);
}

// Stages d0 .. d0+s-1 couple elements inside aligned blocks of n >> d0 = G u elements; with the
// warp-contiguous group mapping a warp owns an aligned block of full_tile / warps elements.
template <int NT> F128_DEV bool pass_is_warp_local(uint32_t n, int d0, uint32_t full_tile)
{
    const bool regular = (full_tile & (full_tile - 1)) == 0 && full_tile >= 256u * (NT / 32); // exact dealing for S <= 3
    return regular && (n >> d0) <= full_tile / (NT / 32);
}
// barrier between two consecutive shared-memory passes: a block barrier unless both are warp-local
template <int NT> F128_DEV void pass_barrier(uint32_t n, int d0_a, int d0_b, uint32_t full_tile)
{
    if (pass_is_warp_local<NT>(n, d0_a, full_tile) && pass_is_warp_local<NT>(n, d0_b, full_tile)) __syncwarp();
    else __syncthreads();
}

template <bool FWD, int NT, int MINB = 512 / NT>
__global__ void __launch_bounds__(NT, MINB)
f128_tile_kernel(Planes data, uint64_t total, uint32_t tile, uint32_t n, uint32_t logn, PassList passes,
                 const Tw4 *__restrict__ tw, bool vec, uint32_t ahead)
{
    constexpr int SMAX = MINB == 3 ? 2 : 3; // three CTAs per SM leave 80 registers: two-stage groups only
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *sre = reinterpret_cast<double2 *>(smem_raw);
    double2 *sim = sre + tile;
    const uint64_t start = uint64_t(blockIdx.x) * tile;
    const uint32_t valid = (total - start < tile) ? uint32_t(total - start) : tile;
    const uint32_t row_off = uint32_t(start & (n - 1));
    Planes g;
#pragma unroll
    for (int i = 0; i < 4; i++) g.p[i] = data.p[i] + start;

    const int last = passes.count - 1;
    if (last == 0) {
        dispatch_pass<FWD, true, true, NT, SMAX>(passes.s[0], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[0], tw, vec);
        return;
    }
    dispatch_pass<FWD, true, false, NT, SMAX>(passes.s[0], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[0], tw, vec);
    // With one or two CTAs per SM the first pass waits for HBM with nothing else to run (FP64-bound kernel, 20 % of a
    // 4096-element tile's time).  Ask L2 for the tile that the CTA taking this one's place will load (`ahead` tiles on =
    // one wave of resident CTAs) while this tile computes; its first pass then finds the data in L2.
    if (ahead) {
        const uint64_t nstart = start + uint64_t(ahead) * tile;
        if (nstart < total) {
            const uint32_t nvalid = (total - nstart < tile) ? uint32_t(total - nstart) : tile;
            for (uint32_t i = threadIdx.x * 16u; i < nvalid; i += NT * 16u) // one 128-byte line per plane and step
#pragma unroll
                for (int pl = 0; pl < 4; pl++) asm volatile("prefetch.global.L2 [%0];" ::"l"(data.p[pl] + nstart + i));
        }
    }
    int prev_d0 = passes.d0[0];
#pragma unroll
    for (int pi = 1; pi < kMaxPasses - 1; pi++) {
        if (pi < last) {
            pass_barrier<NT>(n, prev_d0, passes.d0[pi], tile);
            dispatch_pass<FWD, false, false, NT, SMAX>(passes.s[pi], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[pi], tw, vec);
            prev_d0 = passes.d0[pi];
        }
    }
    // the last pass index is not a compile-time constant: select its parameters without indexing
    int ls = passes.s[1], ld = passes.d0[1];
#pragma unroll
    for (int pi = 2; pi < kMaxPasses; pi++)
        if (pi == last) { ls = passes.s[pi]; ld = passes.d0[pi]; }
    pass_barrier<NT>(n, prev_d0, ld, tile);
    dispatch_pass<FWD, false, true, NT, SMAX>(ls, g, sre, sim, valid, tile, row_off, n, logn, ld, tw, vec);
}

// ---- fwd -> point-wise product with a Fourier-domain operand -> inv in ONE kernel ------
// The negacyclic polynomial product as the reference's own test runs it (src/fft128/mod.rs:2018-2053): Plan::fwd on
// the left operand, lhs <- cplx_mul(lhs, rhs) * factor with the SCALAR double-double product (Scalar::cplx_mul
// :310-326 over mul_f128_f128, f128_ops.rs:395-400 -- not the FMA form the butterflies use), Plan::inv.  As three
// launches the product is a streaming kernel moving 96 B per point; here the tile stays in shared memory between
// the last forward pass and the first inverse pass, so only the right operand is read (32 B per point, from L2 when
// it is shared by the batch).  Same operations in the same order => the bits of the three separate calls.
F128_DEV dd dd_mul_scalar(dd a, dd b)
{
    const double p = __dmul_rn(a.hi, b.hi);
    const double e = __fma_rn(a.hi, b.hi, -p);
    return quick_two_sum(p, __dadd_rn(e, __dadd_rn(__dmul_rn(a.hi, b.lo), __dmul_rn(a.lo, b.hi))));
}

template <int NT, int MINB = 512 / NT>
__global__ void __launch_bounds__(NT, MINB)
f128_fwd_mul_inv_kernel(Planes lhs, Planes rhs, bool rhs_shared, double factor, uint64_t total, uint32_t tile, uint32_t n,
                        uint32_t logn, PassList passes, const Tw4 *__restrict__ tw, bool vec)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *sre = reinterpret_cast<double2 *>(smem_raw);
    double2 *sim = sre + tile;
    const uint64_t start = uint64_t(blockIdx.x) * tile;
    const uint32_t valid = (total - start < tile) ? uint32_t(total - start) : tile;
    const uint32_t row_off = uint32_t(start & (n - 1));
    Planes g;
#pragma unroll
    for (int i = 0; i < 4; i++) g.p[i] = lhs.p[i] + start;
    const int count = passes.count;

    // forward: every pass ends in shared memory
    dispatch_pass<true, true, false, NT, 3>(passes.s[0], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[0], tw, vec);
#pragma unroll 1
    for (int pi = 1; pi < count; pi++) {
        pass_barrier<NT>(n, passes.d0[pi - 1], passes.d0[pi], tile);
        dispatch_pass<true, false, false, NT, 3>(passes.s[pi], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[pi], tw, vec);
    }
    __syncthreads();
    // lhs <- (lhs * rhs) * factor on the tile; element i of the tile is point (row_off + i) mod n of its transform
    {
        const uint64_t roff = rhs_shared ? 0 : start;
        for (uint32_t i = threadIdx.x; i < valid; i += NT) {
            const uint32_t si = swz(i);
            const uint64_t ri = roff + (rhs_shared ? ((row_off + i) & (n - 1)) : i);
            const double2 lre = sre[si], lim = sim[si];
            const dd ar = {lre.x, lre.y}, ai = {lim.x, lim.y};
            const dd br = {__ldg(rhs.p[0] + ri), __ldg(rhs.p[1] + ri)}, bi = {__ldg(rhs.p[2] + ri), __ldg(rhs.p[3] + ri)};
            const dd rr = dd_mul_scalar(ar, br), rim = dd_mul_scalar(ar, bi), ir = dd_mul_scalar(ai, br), ii = dd_mul_scalar(ai, bi);
            const dd pr = dd_sub(rr, ii), pim = dd_add(ir, rim);
            sre[si] = make_double2(__dmul_rn(pr.hi, factor), __dmul_rn(pr.lo, factor));
            sim[si] = make_double2(__dmul_rn(pim.hi, factor), __dmul_rn(pim.lo, factor));
        }
    }
    __syncthreads();
    // inverse: the forward passes in reverse order, the last one writes HBM
#pragma unroll 1
    for (int pi = count - 1; pi >= 1; pi--) {
        dispatch_pass<false, false, false, NT, 3>(passes.s[pi], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[pi], tw, vec);
        pass_barrier<NT>(n, passes.d0[pi], passes.d0[pi - 1], tile);
    }
    dispatch_pass<false, false, true, NT, 3>(passes.s[0], g, sre, sim, valid, tile, row_off, n, logn, passes.d0[0], tw, vec);
}

// stages whose span exceeds a tile: one in-place pass through HBM
template <int S, bool FWD>
__global__ void __launch_bounds__(kThreads, 2)
f128_global_pass(Planes data, uint64_t total, uint32_t n, uint32_t logn, int d0, const Tw4 *__restrict__ tw)
{
    constexpr int G = 1 << S;
    const uint32_t u = n >> (d0 + S);
    const uint64_t groups = total / G;
    for (uint64_t g = uint64_t(blockIdx.x) * kThreads + threadIdx.x; g < groups; g += uint64_t(gridDim.x) * kThreads) {
        const uint64_t hi = g / u;
        const uint32_t lo = uint32_t(g - hi * u);
        const uint64_t base = hi * (uint64_t(G) * u) + lo;
        ddc z[G];
#pragma unroll
        for (int e = 0; e < G; e++) {
            const uint64_t i = base + uint64_t(e) * u;
            z[e] = {{data.p[0][i], data.p[1][i]}, {data.p[2][i], data.p[3][i]}};
        }
        // this thread's NEXT group: its 4 x G lines are requested into L2 while the current group's ~G/2 * S * 94 FP64
        // instructions run (lanes hold consecutive elements, so one lane in sixteen covers a 128-byte line)
        if ((threadIdx.x & 15) == 0) {
            const uint64_t g2 = g + uint64_t(gridDim.x) * kThreads;
            if (g2 < groups) {
                const uint64_t hi2 = g2 / u;
                const uint64_t base2 = hi2 * (uint64_t(G) * u) + (g2 - hi2 * u);
#pragma unroll
                for (int e = 0; e < G; e++)
#pragma unroll
                    for (int pl = 0; pl < 4; pl++) asm volatile("prefetch.global.L2 [%0];" ::"l"(data.p[pl] + base2 + uint64_t(e) * u));
            }
        }
        run_group<S, FWD>(z, tw, uint32_t(base & (n - 1)), logn, d0);
#pragma unroll
        for (int e = 0; e < G; e++) {
            const uint64_t i = base + uint64_t(e) * u;
            data.p[0][i] = z[e].re.hi;
            data.p[1][i] = z[e].re.lo;
            data.p[2][i] = z[e].im.hi;
            data.p[3][i] = z[e].im.lo;
        }
    }
}

template <bool FWD>
cudaError_t launch_global(int S, Planes data, uint64_t total, uint32_t n, uint32_t logn, int d0, const Tw4 *tw,
                          cudaStream_t stream)
{
    uint64_t blocks = (total / (1u << S) + kThreads - 1) / kThreads;
    if (blocks > 148ull * 8) blocks = 148ull * 8;
    const unsigned gsz = unsigned(blocks);
    if (S == 3) f128_global_pass<3, FWD><<<gsz, kThreads, 0, stream>>>(data, total, n, logn, d0, tw);
    else if (S == 2) f128_global_pass<2, FWD><<<gsz, kThreads, 0, stream>>>(data, total, n, logn, d0, tw);
    else f128_global_pass<1, FWD><<<gsz, kThreads, 0, stream>>>(data, total, n, logn, d0, tw);
    count_launch();
    return cudaGetLastError();
}

constexpr uint32_t kF128TileMax = 4096; // 4 planes x 8 B x 4096 = 128 KiB of shared memory

template <bool FWD, int NT, int MINB = 512 / NT>
cudaError_t configure_tile_kernel()
{
    static thread_local int configured_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device == dev) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(f128_tile_kernel<FWD, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         int(kF128TileMax * 4 * sizeof(double)));
    if (e == cudaSuccess) configured_device = dev;
    return e;
}

} // namespace

cudaError_t launch_f128(const cfft_plan *plan, bool inverse, double *re0, double *re1, double *im0, double *im1,
                        uint64_t batch, cudaStream_t stream)
{
    const uint32_t n = uint32_t(plan->n), logn = ilog2(plan->n);
    if (batch == 0) return cudaSuccess;
    const uint64_t total = uint64_t(n) * batch;
    Planes data = {{re0, re1, im0, im1}};
    const Tw4 *tw = reinterpret_cast<const Tw4 *>(plan->d_f128_tw4);

    // Stage split.  n <= 4096: every stage runs in the tile kernel.  Larger n: the widest stages run
    // as HBM passes of three stages each (a 3-stage pass does as much FP64 work as its HBM traffic
    // costs, so it keeps both busy), the rest on sub-blocks of n >> D0 elements in the tile kernel.
    uint32_t tile;
    int D0 = 0; // stages d < D0 run as HBM passes
    static const uint32_t env_tile_max = [] {
        const char *e = getenv("CFFT_B200_F128_TILEMAX");
        const long v = e ? atol(e) : 0;
        return (v == 2048 || v == 4096) ? uint32_t(v) : 0u;
    }();
    // n >= 4096: 2048-element tiles after one more HBM pass (two CTAs per SM) by default, 4096-element tiles (one CTA per SM)
    // when plan->tile_elems == 4096 (the autotuner times both).  Measured with the prefetching HBM pass
    // (profiles/r2g_f128.txt, T FP64 instr/s fwd / inv): n = 4096 13.2 / 13.5 vs 12.4 / 12.9, n = 2^15 13.0 / 13.2 vs 12.3 / 12.7,
    // n = 2^18 12.9 / 12.8 vs 12.3 / 12.5 (n = 8192 and 2^16 end up with 2048-element tiles either way).
    const uint32_t tile_max = env_tile_max ? env_tile_max : ((n >= 4096 && plan->tile_elems == 4096) ? kF128TileMax : 2048u);
    if (n > tile_max) {
        const int over = int(logn - ilog2(tile_max));
        D0 = 3 * ((over + 2) / 3);
        const uint32_t sub = n >> D0;
        tile = sub < 2048 ? 2048 : sub; // whole sub-blocks, 2048 or 4096 elements
        if (uint64_t(tile) > uint64_t(n) * batch) tile = sub;
    } else if (n >= 2048 && !(plan->tile_elems > n)) {
        tile = n;
    } else {
        uint64_t rows = (plan->tile_elems ? plan->tile_elems : 2048u) / n;
        if (rows < 1) rows = 1;
        if (rows > batch) rows = batch;
        tile = uint32_t(rows * n);
    }

    int gd0[16], gs[16], gcount = 0;
    for (int d = 0; d < D0; d += 3) {
        gd0[gcount] = d;
        gs[gcount++] = 3;
    }
    // tile passes: ceil(k / 3) passes with the stages spread evenly (3,3,2,2 rather than 3,3,3,1)
    static const int env_smax = [] { const char *e = getenv("CFFT_B200_F128_SMAX"); return e ? atoi(e) : 0; }();
    const int smax = (env_smax == 2 || env_smax == 3) ? env_smax : (plan->f128_smax == 2 ? 2 : 3);
    PassList passes;
    passes.count = 0;
    {
        const int k = int(logn) - D0;
        const int np = (k + smax - 1) / smax, lo = k / np, extra = k % np;
        for (int i = 0, d = D0; i < np; i++) {
            const int sz = lo + (i < extra ? 1 : 0);
            passes.d0[passes.count] = d;
            passes.s[passes.count++] = sz;
            d += sz;
        }
    }
    const size_t smem = size_t(tile) * 4 * sizeof(double);
    // 128-bit HBM accesses in the stride-1 pass need 16-byte aligned planes
    const bool vec = ((reinterpret_cast<uintptr_t>(re0) | reinterpret_cast<uintptr_t>(re1) | reinterpret_cast<uintptr_t>(im0) |
                       reinterpret_cast<uintptr_t>(im1)) & 15) == 0;
    const unsigned tiles = unsigned((total + tile - 1) / tile);
    cudaError_t e;
    // L2 prefetch distance of the tile kernel = one wave of resident CTAs (SMs x CTAs per SM by shared memory);
    // CFFT_B200_F128_PREFETCH=0 turns it off, =k sets the distance in waves
    static const int env_pf = [] { const char *e = getenv("CFFT_B200_F128_PREFETCH"); return e ? atoi(e) : 1; }();
    uint32_t ahead = 0;
    if (env_pf > 0 && tile > 2048) { // measured (profiles/r2c_f128_prefetch.txt): +2 % at n = 4096 and n = 2^15, nothing (or -1 %) with two CTAs per SM
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, plan->device);
        const uint32_t per_sm = tile > 2048 ? 1u : (smax == 2 ? 3u : 2u);
        ahead = uint32_t(sms) * per_sm * uint32_t(env_pf);
    }

    if (!inverse) {
        for (int i = 0; i < gcount; i++)
            if ((e = launch_global<true>(gs[i], data, total, n, logn, gd0[i], tw, stream)) != cudaSuccess) return e;
        // 8 elements per thread and pass: a 4096-element tile gets 512 threads (16 warps per SM)
        if (tile > 2048) {
            if ((e = configure_tile_kernel<true, 512>()) != cudaSuccess) return e;
            f128_tile_kernel<true, 512><<<tiles, 512, smem, stream>>>(data, total, tile, n, logn, passes, tw, vec, ahead);
        } else if (smax == 2) {
            if ((e = configure_tile_kernel<true, 256, 3>()) != cudaSuccess) return e;
            f128_tile_kernel<true, 256, 3><<<tiles, 256, smem, stream>>>(data, total, tile, n, logn, passes, tw, vec, ahead);
        } else {
            if ((e = configure_tile_kernel<true, 256>()) != cudaSuccess) return e;
            f128_tile_kernel<true, 256><<<tiles, 256, smem, stream>>>(data, total, tile, n, logn, passes, tw, vec, ahead);
        }
        count_launch();
        return cudaGetLastError();
    }
    // inverse: same passes in reverse order
    PassList rev;
    rev.count = passes.count;
    for (int i = 0; i < passes.count; i++) {
        rev.d0[i] = passes.d0[passes.count - 1 - i];
        rev.s[i] = passes.s[passes.count - 1 - i];
    }
    if (tile > 2048) {
        if ((e = configure_tile_kernel<false, 512>()) != cudaSuccess) return e;
        f128_tile_kernel<false, 512><<<tiles, 512, smem, stream>>>(data, total, tile, n, logn, rev, tw, vec, ahead);
    } else if (smax == 2) {
        if ((e = configure_tile_kernel<false, 256, 3>()) != cudaSuccess) return e;
        f128_tile_kernel<false, 256, 3><<<tiles, 256, smem, stream>>>(data, total, tile, n, logn, rev, tw, vec, ahead);
    } else {
        if ((e = configure_tile_kernel<false, 256>()) != cudaSuccess) return e;
        f128_tile_kernel<false, 256><<<tiles, 256, smem, stream>>>(data, total, tile, n, logn, rev, tw, vec, ahead);
    }
    count_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    for (int i = gcount - 1; i >= 0; i--)
        if ((e = launch_global<false>(gs[i], data, total, n, logn, gd0[i], tw, stream)) != cudaSuccess) return e;
    return cudaSuccess;
}

bool f128_fused_mul_kernel_available(const cfft_plan *plan) { return plan->n >= 32 && plan->n <= kF128TileMax; }

// lhs <- inv( (fwd(lhs) * rhs) * factor ) on `batch` rows; rhs is in the Fourier domain (this plan's bit-reversed
// order), one row shared by the batch or one row per transform.  n <= 4096: one kernel; larger n: the three launches.
cudaError_t launch_f128_fwd_mul_inv(const cfft_plan *plan, double *l_re0, double *l_re1, double *l_im0, double *l_im1,
                                    const double *r_re0, const double *r_re1, const double *r_im0, const double *r_im1,
                                    bool rhs_shared, double factor, uint64_t batch, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    const uint32_t n = uint32_t(plan->n), logn = ilog2(plan->n);
    const bool composed = !f128_fused_mul_kernel_available(plan) || getenv("CFFT_B200_FUSED_MUL_COMPOSED") != nullptr;
    if (composed) {
        cudaError_t e = launch_f128(plan, false, l_re0, l_re1, l_im0, l_im1, batch, stream);
        if (e == cudaSuccess)
            e = launch_f128_cplx_mul_scale_rows(l_re0, l_re1, l_im0, l_im1, r_re0, r_re1, r_im0, r_im1, rhs_shared ? n : 0, factor,
                                                uint64_t(n) * batch, stream);
        if (e == cudaSuccess) e = launch_f128(plan, true, l_re0, l_re1, l_im0, l_im1, batch, stream);
        return e;
    }
    const uint64_t total = uint64_t(n) * batch;
    uint32_t tile = n;
    if (n < 2048) {
        uint64_t rows = 2048u / n;
        if (rows > batch) rows = batch;
        tile = uint32_t(rows * n);
    }
    PassList passes;
    passes.count = 0;
    {
        const int k = int(logn), np = (k + 2) / 3, lo = k / np, extra = k % np;
        for (int i = 0, d = 0; i < np; i++) {
            const int sz = lo + (i < extra ? 1 : 0);
            passes.d0[passes.count] = d;
            passes.s[passes.count++] = sz;
            d += sz;
        }
    }
    const size_t smem = size_t(tile) * 4 * sizeof(double);
    const bool vec = ((reinterpret_cast<uintptr_t>(l_re0) | reinterpret_cast<uintptr_t>(l_re1) | reinterpret_cast<uintptr_t>(l_im0) |
                       reinterpret_cast<uintptr_t>(l_im1)) & 15) == 0;
    Planes lhs = {{l_re0, l_re1, l_im0, l_im1}};
    Planes rhs = {{const_cast<double *>(r_re0), const_cast<double *>(r_re1), const_cast<double *>(r_im0), const_cast<double *>(r_im1)}};
    const Tw4 *tw = reinterpret_cast<const Tw4 *>(plan->d_f128_tw4);
    const unsigned tiles = unsigned((total + tile - 1) / tile);
    static thread_local int configured_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device != dev) {
        cudaError_t e = cudaFuncSetAttribute(f128_fwd_mul_inv_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(kF128TileMax * 4 * sizeof(double)));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(f128_fwd_mul_inv_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     int(kF128TileMax * 4 * sizeof(double)));
        if (e != cudaSuccess) return e;
        configured_device = dev;
    }
    if (tile > 2048)
        f128_fwd_mul_inv_kernel<512><<<tiles, 512, smem, stream>>>(lhs, rhs, rhs_shared, factor, total, tile, n, logn, passes, tw, vec);
    else
        f128_fwd_mul_inv_kernel<256><<<tiles, 256, smem, stream>>>(lhs, rhs, rhs_shared, factor, total, tile, n, logn, passes, tw, vec);
    count_launch();
    return cudaGetLastError();
}

} // namespace cfft

// c64_fast_kernels.cuh -- the register-resident c64 kernels for plans with base (Dif16, 256), as templates shared by
// c64_fast.cu (plain c64 rows) and c64_poly.cu (integer polynomials in / out).  See c64_fast.cu for the
// design notes; every kernel reads and writes global memory only through the row accessors of c64_dev.cuh.
#pragma once
#include <mutex>
#include <set>
#include <utility>

#include "c64_dev.cuh"
#include "plan.h"

namespace cfft {
namespace fastk {
using namespace dev;

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// One unordered level of span NCUR on the 16 register values of a thread: B = 16/R butterflies,
// butterfly j covers positions base_j + m*k.  Twiddles planar: tw[(k-1)*m + p].
// G_IN / G_OUT: the level reads / writes the row in global memory through the accessor `io` (c64_dev.cuh: plain c64 or a
// polynomial with the fold / twist / rounding fused in), else the shared-memory tile ssrc / sdst.
template <int R, int NCUR, int TPR, bool FWD, bool G_IN, bool G_OUT, class Io>
__device__ __forceinline__ void level_io(const Io &io, const c64 *__restrict__ ssrc, c64 *__restrict__ sdst,
                                         const c64 *__restrict__ tw, int t, c64 (&v)[16])
{
    constexpr int B = 16 / R, m = NCUR / R;
    int base[B], p[B];
#pragma unroll
    for (int j = 0; j < B; j++) {
        const int b = t + TPR * j;
        const int blk = b / m;
        p[j] = b - blk * m;
        base[j] = blk * NCUR + p[j];
    }
#pragma unroll
    for (int j = 0; j < B; j++)
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int pos = base[j] + m * (FWD ? k : brev_c<R>(k));
            v[j * R + k] = G_IN ? io.ld(pos) : ssrc[pos];
        }
    if (!G_IN && !G_OUT) __syncthreads(); // everyone has read before anyone overwrites (in place)
#pragma unroll
    for (int j = 0; j < B; j++) {
        c64 *x = &v[j * R];
        if (!FWD) {
#pragma unroll
            for (int k = 1; k < R; k++) x[k] = cmul(ld_tw(tw + (k - 1) * m + p[j]), x[k]);
        }
        bfR<R, FWD>(x);
        if (FWD) {
#pragma unroll
            for (int k = 1; k < R; k++) x[k] = cmul(ld_tw(tw + (k - 1) * m + p[j]), x[k]);
        }
    }
#pragma unroll
    for (int j = 0; j < B; j++)
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int pos = base[j] + m * (FWD ? brev_c<R>(k) : k);
            if (G_OUT) io.st(pos, v[j * R + k]);
            else sdst[pos] = v[j * R + k];
        }
}
// pointer form: src / dst are global when G_IN / G_OUT, shared memory otherwise
template <int R, int NCUR, int TPR, bool FWD, bool G_IN, bool G_OUT>
__device__ __forceinline__ void level(const c64 *__restrict__ src, c64 *__restrict__ dst, const c64 *__restrict__ tw, int t, c64 (&v)[16])
{
    level_io<R, NCUR, TPR, FWD, G_IN, G_OUT>(plain_row(G_IN ? src : nullptr, G_OUT ? dst : nullptr), src, dst, tw, t, v);
}

// Two consecutive unordered levels (radix 8 on span NCUR, then radix 2 on span NCUR/8) fused in
// registers: a thread's two radix-8 butterflies sit at p = t and t + NCUR/16, which is exactly the
// pair the radix-2 level combines inside every chunk, so no exchange is needed between them.
// Arithmetic per butterfly is unchanged (src/unordered.rs:24-43, 98-219).
template <int NCUR, int TPR, bool FWD, class Io>
__device__ __forceinline__ void level_8x2_io(const Io &io, c64 *__restrict__ s, const c64 *__restrict__ tw8, const c64 *__restrict__ tw2,
                                             int t, c64 (&v)[16])
{
    static_assert(TPR * 16 == NCUR, "thread t owns p = t and p = t + NCUR/16");
    constexpr int m = NCUR / 8;  // 8 chunks of m after the radix-8 level
    constexpr int h = NCUR / 16; // half a chunk = span of the radix-2 level's butterflies
    if (FWD) {
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 8; k++) v[j * 8 + k] = io.ld(t + h * j + m * k);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            c64 *x = &v[j * 8];
            bf8<true>(x);
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(tw8 + (k - 1) * m + t + h * j), x[k]);
        }
        const c64 w = ld_tw(tw2 + t);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const c64 a = v[k], b = v[8 + k];
            const int chunk = m * brev_c<8>(k);
            s[chunk + t] = cadd(a, b);              // fwd_butterfly_x2
            s[chunk + h + t] = cmul(w, csub(a, b));
        }
    } else {
        const c64 w = ld_tw(tw2 + t);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int chunk = m * brev_c<8>(k);
            const c64 a = s[chunk + t], b = cmul(w, s[chunk + h + t]); // inv_butterfly_x2
            v[k] = cadd(a, b);
            v[8 + k] = csub(a, b);
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            c64 *x = &v[j * 8];
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(tw8 + (k - 1) * m + t + h * j), x[k]);
            bf8<false>(x);
#pragma unroll
            for (int k = 0; k < 8; k++) io.st(t + h * j + m * k, x[k]);
        }
    }
}
template <int NCUR, int TPR, bool FWD>
__device__ __forceinline__ void level_8x2(const c64 *__restrict__ gsrc, c64 *__restrict__ gdst, c64 *__restrict__ s,
                                          const c64 *__restrict__ tw8, const c64 *__restrict__ tw2, int t, c64 (&v)[16])
{
    level_8x2_io<NCUR, TPR, FWD>(plain_row(gsrc, gdst), s, tw8, tw2, t, v);
}

struct FastTables {
    const c64 *top1; // planar (R1-1) x n/R1
    const c64 *top2; // planar (R2-1) x n/(R1 R2)
    const c64 *base; // planar half of init_wt(16, 256): w[p + 16 k]
};

// the planar tables of a (Dif16, 256) plan for one direction (0 fwd, 1 inv)
inline FastTables fast_tables(const cfft_plan *plan, int dir)
{
    const c64 *base = plan->d_fast_tw[dir];
    FastTables tb;
    tb.top1 = plan->fast_levels.size() > 0 ? base + plan->fast_levels[0].off : base;
    tb.top2 = plan->fast_levels.size() > 1 ? base + plan->fast_levels[1].off : base;
    tb.base = base + plan->fast_base_off;
    return tb;
}

// opt a kernel in to more than 48 KiB of dynamic shared memory, once per (kernel, device)
template <class K> inline cudaError_t allow_smem(K kernel, size_t smem)
{
    if (smem <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::set<std::pair<const void *, int>> done; // K is only the TYPE of the kernel pointer: key on its value
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const std::pair<const void *, int> key(reinterpret_cast<const void *>(kernel), dev);
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(key)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e == cudaSuccess) done.insert(key);
    return e;
}

template <int N> struct FastCfg {
    static constexpr int TPR = N / 16;                  // threads per transform
    static constexpr int NT = TPR < 128 ? 128 : TPR;    // threads per CTA
    static constexpr int ROWS = NT / TPR;               // transforms per CTA
    static constexpr int MINB = (NT <= 128) ? 4 : (NT <= 256 ? 2 : 1);
};

// STD = true: standard-order ("ordered") in / out.  X_i sits in the unordered layout at chunk
// c = bitrev_L(i mod M), offset i / M (M = N / 256 chunks, src/unordered.rs:1046-1051).  Round 1 added one more
// shared-memory exchange for the un-permutation; since round 2 it rides on the base FFT's own 16 x 16 transpose (the two
// radix-16 passes of a chunk need not run on the same half-warp), so the standard-order kernel moves each element through
// shared memory exactly as often as the unordered one and HBM still sees consecutive i only: n = 2048 6.10 -> 6.91 TB/s,
// n = 8192 3.9 -> 4.1-4.4 (profiles/r2l_std_one_exchange_ab.txt).
// PIN / POUT: the rows in global memory are integer polynomials on the input / output side (c64_dev.cuh, RowIo).
template <int N, int R1, int R2, bool FWD, bool STD = false, bool PIN = false, bool POUT = false>
__global__ void __launch_bounds__(FastCfg<N>::NT, FastCfg<N>::MINB)
c64_fast_b256_kernel(BatchIo<PIN, POUT> bio, uint64_t batch, FastTables tb)
{
    static_assert(N == 256 * R1 * R2, "n = 256 * R1 * R2");
    static_assert(!STD || N >= 2048, "standard-order variant: M = N / 256 >= 8 chunks");
    using Cfg = FastCfg<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int row = threadIdx.x / Cfg::TPR, t = threadIdx.x % Cfg::TPR;
    const uint64_t grow = uint64_t(blockIdx.x) * Cfg::ROWS + row;
    // rows are whole-warp aligned (TPR >= 16 and ROWS * TPR = NT), but a CTA may own fewer rows at
    // the tail; inactive rows still take part in the block barriers below.
    const bool active = grow < batch;
    const RowIo<PIN, POUT> io = bio.row(active ? grow : 0);
    c64 *s = reinterpret_cast<c64 *>(smem_raw) + row * N;
    c64 v[16];
    constexpr int N2 = N / R1;      // span of the second level
    const int blk = t / 16, lane16 = t % 16;

    constexpr int M = N / 256, LOGM = (M == 8 ? 3 : (M == 16 ? 4 : 5));
    const int sw = STD ? int(__brev(unsigned(blk)) >> (32 - LOGM)) & 7 : 0; // (lo & 7) of this thread's chunk
    constexpr bool FUSED = (R1 == 8 && R2 == 2); // both levels in registers, see level_8x2
    // Large transforms leave one or two CTAs per SM, so a CTA's first loads wait for HBM with little else to run: once
    // its own loads are done it asks L2 for the row the CTA that takes its place will start with (`ahead` rows on = one
    // wave of resident CTAs), N / 8 lines of 128 bytes = two requests per thread.
    auto prefetch_successor = [&] {
        if (!PIN && bio.ahead && grow + bio.ahead < batch) {
            const c64 *nx = bio.in + (grow + bio.ahead) * bio.row_in;
#pragma unroll
            for (int i = 0; i < 2; i++) prefetch_l2(nx + (t + Cfg::TPR * i) * 8);
        }
    };
    if (FWD) {
        if (FUSED) {
            if (active) level_8x2_io<N, Cfg::TPR, true>(io, s, tb.top1, tb.top2, t, v);
            prefetch_successor();
            __syncthreads();
        } else if (R1 > 1) {
            if (active) level_io<R1, N, Cfg::TPR, true, true, false>(io, nullptr, s, tb.top1, t, v);
            prefetch_successor();
            __syncthreads();
        }
        if (R2 > 1 && !FUSED) {
            level<R2, N2, Cfg::TPR, true, false, false>(s, s, tb.top2, t, v);
            __syncthreads();
        }
        if (STD) {
            // The un-permutation folded into the base FFT's OWN transpose: pass 1 (radix 16 on x[p + 16k]) runs with half-warp =
            // chunk, lane = p; pass 2 (radix 16 on y[j + 16k']) runs with lanes on W = min(M, 16) CONSECUTIVE lo, so that a lane
            // group finishes with X[hi = j + 16k''] of consecutive standard indices i = hi M + lo and stores full lines straight
            // from registers -- no natural-order write-back, no transposing read.  The transposed block of chunk c is XORed
            // with (lo & 7) in its low three bits: lanes on consecutive p (write) and lanes on consecutive lo (read) both land in
            // eight different 16-byte bank groups.
            c64 *sb = s + blk * 256;
            const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = sb[lane16 + 16 * k];
            bf16<true>(v);
#pragma unroll
            for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tb.base + lane16 + 16 * k), v[k]);
            __syncwarp(hmask);
#pragma unroll
            for (int k = 0; k < 16; k++) sb[16 * lane16 + ((k ^ lane16) ^ sw)] = v[k];
            __syncthreads();
            constexpr int W = M < 16 ? M : 16;
            const int lo = (t % W) + W * (t / (16 * W)), j = (t / W) % 16;
            const c64 *sc = s + int(__brev(unsigned(lo)) >> (32 - LOGM)) * 256;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = sc[16 * k + ((j ^ k) ^ (lo & 7))];
            bf16<true>(v);
            if (active) {
#pragma unroll
                for (int k = 0; k < 16; k++) io.st((j + 16 * k) * M + lo, v[k]);
            }
        } else if (active) {
            if (R1 > 1) base256_io<true, false, true>(io, blk * 256, s + blk * 256, s + blk * 256, nullptr, tb.base, lane16, v);
            else base256_io<true, true, true>(io, blk * 256, nullptr, s + blk * 256, nullptr, tb.base, lane16, v);
        }
    } else {
        if (STD) {
            // mirror image: pass 1 on standard-order input with lanes on consecutive lo, pass 2 on the thread's own chunk
            constexpr int W = M < 16 ? M : 16;
            const int lo = (t % W) + W * (t / (16 * W)), p = (t / W) % 16;
            c64 *sc = s + int(__brev(unsigned(lo)) >> (32 - LOGM)) * 256;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = io.ld((p + 16 * k) * M + lo);
            bf16<false>(v);
#pragma unroll
            for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tb.base + p + 16 * k), v[k]);
#pragma unroll
            for (int k = 0; k < 16; k++) sc[16 * p + ((k ^ p) ^ (lo & 7))] = v[k];
            __syncthreads();
            c64 *sb = s + blk * 256;
            const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = sb[16 * k + ((lane16 ^ k) ^ sw)];
            bf16<false>(v);
            __syncwarp(hmask); // the half-warp has consumed the transposed block before it is overwritten in natural order
#pragma unroll
            for (int k = 0; k < 16; k++) sb[lane16 + 16 * k] = v[k];
        } else if (active) {
            if (R1 > 1) base256_io<false, true, false>(io, blk * 256, nullptr, s + blk * 256, s + blk * 256, tb.base, lane16, v);
            else base256_io<false, true, true>(io, blk * 256, nullptr, s + blk * 256, nullptr, tb.base, lane16, v);
        }
        if (R1 > 1) prefetch_successor();
        if (FUSED) {
            __syncthreads();
            if (active) level_8x2_io<N, Cfg::TPR, false>(io, s, tb.top1, tb.top2, t, v);
        } else {
            if (R2 > 1) {
                __syncthreads();
                level<R2, N2, Cfg::TPR, false, false, false>(s, s, tb.top2, t, v);
            }
            if (R1 > 1) {
                __syncthreads();
                if (active) level_io<R1, N, Cfg::TPR, false, false, true>(io, s, nullptr, tb.top1, t, v);
            }
        }
    }
}

// ---- fwd -> point-wise multiply-accumulate -> inv in ONE kernel ----------
// out[r] = inv( sum_k fwd(a[r][k]) (.) b[k] ): the external-product / convolution step of a caller that keeps
// its data on the GPU.  As three library calls per term (cfft_c64_fwd, cfft_c64_mul_[add_]assign, cfft_c64_inv)
// a K = 1 product moves 7 x 16 n bytes through HBM; here it moves 3 x 16 n (2 x 16 n when b is shared by the
// batch and therefore L2-resident), and K terms cost (2K + 1) x 16 n instead of (6K + 1) x 16 n.  The forward
// transform of a term ends with every thread holding 16 Fourier coefficients in registers at exactly the
// positions the inverse base FFT starts from (blk * 256 + lane16 + 16 j), so the product and the running sum
// never leave the SM: the sum of K > 1 terms lives in 16 more c64 registers per thread.  Butterflies, twiddles
// and the order of every rounded operation are those of the separate kernels, the product is num_complex's (no FMA), terms are added in the order k = 0, 1, ... =>
// bit-identical to the composition of the library calls (tests/test_gpu_c64.py).
template <bool FWD>
__device__ __forceinline__ void base256_core(c64 *__restrict__ sm_blk, const c64 *__restrict__ tw_planar, int lane16, c64 (&v)[16])
{
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16);
    bf16<FWD>(v);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw_planar + lane16 + 16 * k), v[k]);
    __syncwarp(hmask);
#pragma unroll
    for (int k = 0; k < 16; k++) sm_blk[16 * lane16 + (k ^ lane16)] = v[k];
    __syncwarp(hmask);
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = sm_blk[16 * k + (lane16 ^ k)];
    bf16<FWD>(v);
}


// MULTI = false: exactly one term, nothing to accumulate: the register budget and occupancy of the plain kernel.
// MULTI = true: the running sum of the terms lives in 16 more c64 registers per thread (168 registers, three CTAs
// of 128 threads per SM); a shared-memory accumulator was measured first and lost (the L1 / shared-memory pipe is
// already the busiest unit of the plain kernel, and a second tile per CTA leaves too little L1 for the twiddles).
template <int N, bool MULTI> struct FusedMulCfg {
    static constexpr int NT = FastCfg<N>::NT;
    static constexpr int MINB = !MULTI ? FastCfg<N>::MINB : (NT <= 128 ? 3 : 1);
};

// PIN: the terms a[r][k] are integer polynomials (2N coefficients each); POUT: so is out[r] (c64_dev.cuh, RowIo) -- the
// whole negacyclic product step of a caller in one kernel, integers in, integers out.
template <int N, int R1, int R2, bool MULTI, bool PIN = false, bool POUT = false>
__global__ void __launch_bounds__(FusedMulCfg<N, MULTI>::NT, FusedMulCfg<N, MULTI>::MINB)
c64_fwd_mul_inv_kernel(BatchIo<PIN, false> ain, const c64 *__restrict__ b, BatchIo<false, POUT> oout, uint64_t batch,
                       uint32_t kterms, uint64_t b_row_stride, FastTables tf, FastTables ti, uint32_t flags)
{
    static_assert(N == 256 * R1 * R2, "n = 256 * R1 * R2");
    using Cfg = FastCfg<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int row = threadIdx.x / Cfg::TPR, t = threadIdx.x % Cfg::TPR;
    const uint64_t grow = uint64_t(blockIdx.x) * Cfg::ROWS + row;
    const bool active = grow < batch; // inactive rows compute on row 0's data and store nothing
    const uint64_t r = active ? grow : 0;
    const c64 *gb = b + r * b_row_stride;
    const RowIo<false, POUT> io_out = oout.row(r);
    c64 *s = reinterpret_cast<c64 *>(smem_raw) + row * N;
    c64 v[16];
    c64 acc[MULTI ? 16 : 1];
    constexpr int N2 = N / R1;
    constexpr bool FUSED = (R1 == 8 && R2 == 2);
    const int blk = t / 16, lane16 = t % 16;
    c64 *sb = s + blk * 256;

    // the 16 n bytes of b this row's first term needs: request them now, they are consumed last
    // (a row's threads cover its n c64 = n / 8 lines of 128 bytes with 2 requests each)
    if (flags & 1) {
#pragma unroll
        for (int i = 0; i < 2; i++) prefetch_l2(gb + (t + Cfg::TPR * i) * 8);
    }

    const uint32_t kt = MULTI ? kterms : 1;
    for (uint32_t k = 0; k < kt; k++) {
        const RowIo<PIN, false> io = ain.row(r * kterms + k); // rows of `a` are one term apart (fwd_mul_add: kterms = terms per row)
        if (MULTI && k + 1 < kt && (flags & 2)) { // next term's input and multiplier on their way to L2 while this term computes
#pragma unroll
            for (int i = 0; i < 2; i++) {
                if (PIN) prefetch_l2(io.pin + 2 * N + (t + Cfg::TPR * i) * 16);
                else prefetch_l2(io.in + N + (t + Cfg::TPR * i) * 8);
                prefetch_l2(gb + uint64_t(k + 1) * N + (t + Cfg::TPR * i) * 8);
            }
        }
        if (MULTI && k > 0 && R1 > 1) __syncthreads(); // the previous term's base FFTs have consumed the tile
        if (FUSED) {
            level_8x2_io<N, Cfg::TPR, true>(io, s, tf.top1, tf.top2, t, v);
        } else if (R1 > 1) {
            level_io<R1, N, Cfg::TPR, true, true, false>(io, nullptr, s, tf.top1, t, v);
        }
        // own loads of the row's first term are done: ask L2 for the first term (and, in a chained launch, the partial sum)
        // of the row the CTA taking this one's place will start with -- flags >> 8 rows on, one wave of resident CTAs
        if (R1 > 1 && k == 0 && !PIN && (flags >> 8) && grow + (flags >> 8) < batch) {
            const uint64_t r2 = grow + (flags >> 8);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                prefetch_l2(ain.in + r2 * kterms * ain.row_in + (t + Cfg::TPR * i) * 8);
                if (!POUT && (flags & 8)) prefetch_l2(oout.out + r2 * oout.row_out + (t + Cfg::TPR * i) * 8);
            }
        }
        if (R1 > 1) __syncthreads();
        if (R2 > 1 && !FUSED) {
            level<R2, N2, Cfg::TPR, true, false, false>(s, s, tf.top2, t, v);
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = R1 > 1 ? sb[lane16 + 16 * j] : io.ld(blk * 256 + lane16 + 16 * j);
        base256_core<true>(sb, tf.base, lane16, v);
        const c64 *bk = gb + uint64_t(k) * N + blk * 256 + lane16;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const c64 p = cmul_nc(v[j], ld_stream(bk + 16 * j));
            if (!MULTI) v[j] = p;
            else acc[j] = k == 0 ? p : cadd(acc[j], p);
        }
    }
    if (!MULTI) {
        // chained launches (one term each, for sizes without the accumulating kernel): bit 3 adds the Fourier-domain
        // partial sum the previous launch left in `out`, bit 4 stores the new partial sum instead of inverting it
        // (plain c64 `out` only: the polynomial entry points never chain)
        if (!POUT) {
            c64 *gs = io_out.out + blk * 256 + lane16;
            if (flags & 8) {
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = cadd(ld_stream(gs + 16 * j), v[j]);
            }
            if (flags & 16) {
                if (active) {
#pragma unroll
                    for (int j = 0; j < 16; j++) st_stream(gs + 16 * j, v[j]);
                }
                return;
            }
        }
    }
    if (MULTI) {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = acc[j];
    }

    base256_core<false>(sb, ti.base, lane16, v);
    if (R1 == 1) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 16; j++) io_out.st(blk * 256 + lane16 + 16 * j, v[j]);
        }
        return;
    }
    __syncwarp(0xFFFFu << (threadIdx.x & 16));
#pragma unroll
    for (int j = 0; j < 16; j++) sb[lane16 + 16 * j] = v[j];
    if (FUSED) {
        __syncthreads();
        if (active) level_8x2_io<N, Cfg::TPR, false>(io_out, s, ti.top1, ti.top2, t, v);
    } else {
        if (R2 > 1) {
            __syncthreads();
            level<R2, N2, Cfg::TPR, false, false, false>(s, s, ti.top2, t, v);
        }
        __syncthreads();
        if (active) level_io<R1, N, Cfg::TPR, false, false, true>(io_out, s, nullptr, ti.top1, t, v);
    }
}

// ---- the same step with TWO outputs per row: out[r][o] = inv( sum_k fwd(a[r][k]) (.) b[r][k][o] ), o < 2 -------------------
// The GLWE external product of a caller such as TFHE-rs (GLWE dimension 1): every decomposed term a[r][k] feeds BOTH output
// polynomials, each against its own row of the Fourier-domain key.  With cfft_c64_fwd_mul_inv once per output the k forward
// transforms run twice; here each runs once and its 16 coefficients per thread are multiplied into two running sums
// (2 x 16 c64 registers), then the two inverse transforms go through the same shared-memory tile one after the other.
// 2 k + 2 transforms per row instead of 2 k + 2 + 2 k.  Arithmetic per output exactly as c64_fwd_mul_inv_kernel (num_complex
// product, terms added in order) => bit-identical to calling that kernel per output.
template <int N, int R1, int R2>
__global__ void __launch_bounds__(FastCfg<N>::NT, 2)
c64_fwd_mul_inv2_kernel(const c64 *__restrict__ a, const c64 *__restrict__ b, c64 *__restrict__ out, uint64_t batch, uint32_t kterms,
                        uint64_t b_row_stride, FastTables tf, FastTables ti, uint32_t ahead)
{
    static_assert(N == 256 * R1 * R2, "n = 256 * R1 * R2");
    using Cfg = FastCfg<N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int row = threadIdx.x / Cfg::TPR, t = threadIdx.x % Cfg::TPR;
    const uint64_t grow = uint64_t(blockIdx.x) * Cfg::ROWS + row;
    const bool active = grow < batch; // inactive rows compute on row 0's data and store nothing
    const uint64_t r = active ? grow : 0;
    const c64 *gb = b + r * b_row_stride; // [kterms][2][N]
    c64 *s = reinterpret_cast<c64 *>(smem_raw) + row * N;
    c64 v[16], acc0[16], acc1[16];
    constexpr int N2 = N / R1;
    constexpr bool FUSED = (R1 == 8 && R2 == 2);
    const int blk = t / 16, lane16 = t % 16;
    c64 *sb = s + blk * 256;

    for (uint32_t k = 0; k < kterms; k++) {
        const PlainRow io = plain_row(a + (r * kterms + k) * N, nullptr);
        if (k + 1 < kterms) { // next term's input and both multiplier rows on their way to L2 while this term computes
#pragma unroll
            for (int i = 0; i < 2; i++) {
                prefetch_l2(io.in + N + (t + Cfg::TPR * i) * 8);
                prefetch_l2(gb + uint64_t(k + 1) * 2 * N + (t + Cfg::TPR * i) * 8);
                prefetch_l2(gb + uint64_t(k + 1) * 2 * N + N + (t + Cfg::TPR * i) * 8);
            }
        }
        if (k > 0 && R1 > 1) __syncthreads(); // the previous term's base FFTs have consumed the tile
        if (FUSED) {
            level_8x2_io<N, Cfg::TPR, true>(io, s, tf.top1, tf.top2, t, v);
        } else if (R1 > 1) {
            level_io<R1, N, Cfg::TPR, true, true, false>(io, nullptr, s, tf.top1, t, v);
        }
        if (R1 > 1 && k == 0 && ahead && grow + ahead < batch) { // the row of the CTA that takes this one's place
#pragma unroll
            for (int i = 0; i < 2; i++) prefetch_l2(a + (grow + ahead) * kterms * N + (t + Cfg::TPR * i) * 8);
        }
        if (R1 > 1) __syncthreads();
        if (R2 > 1 && !FUSED) {
            level<R2, N2, Cfg::TPR, true, false, false>(s, s, tf.top2, t, v);
            __syncthreads();
        }
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = R1 > 1 ? sb[lane16 + 16 * j] : io.ld(blk * 256 + lane16 + 16 * j);
        base256_core<true>(sb, tf.base, lane16, v);
        const c64 *bk = gb + uint64_t(k) * 2 * N + blk * 256 + lane16;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const c64 p0 = cmul_nc(v[j], ld_stream(bk + 16 * j)), p1 = cmul_nc(v[j], ld_stream(bk + N + 16 * j));
            acc0[j] = k == 0 ? p0 : cadd(acc0[j], p0);
            acc1[j] = k == 0 ? p1 : cadd(acc1[j], p1);
        }
    }
#pragma unroll
    for (int o = 0; o < 2; o++) {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = o == 0 ? acc0[j] : acc1[j];
        if (o == 1 && R1 > 1) __syncthreads(); // the first output's levels have consumed the tile
        base256_core<false>(sb, ti.base, lane16, v);
        const PlainRow io_out = plain_row(nullptr, out + (r * 2 + o) * N);
        if (R1 == 1) {
            if (active) {
#pragma unroll
                for (int j = 0; j < 16; j++) io_out.st(blk * 256 + lane16 + 16 * j, v[j]);
            }
            continue;
        }
        __syncwarp(0xFFFFu << (threadIdx.x & 16));
#pragma unroll
        for (int j = 0; j < 16; j++) sb[lane16 + 16 * j] = v[j];
        if (FUSED) {
            __syncthreads();
            if (active) level_8x2_io<N, Cfg::TPR, false>(io_out, s, ti.top1, ti.top2, t, v);
        } else {
            if (R2 > 1) {
                __syncthreads();
                level<R2, N2, Cfg::TPR, false, false, false>(s, s, ti.top2, t, v);
            }
            __syncthreads();
            if (active) level_io<R1, N, Cfg::TPR, false, false, true>(io_out, s, nullptr, ti.top1, t, v);
        }
    }
}

} // namespace fastk
} // namespace cfft

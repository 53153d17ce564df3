// c64_column.cu -- the unordered levels of large transforms (n >= 2^14) for plans with base
// (Dif16, 256): groups of up to three consecutive radix-2/4/8 levels
// (fwd_process_x* / inv_process_x*, src/unordered.rs:222-293) executed in one HBM pass.
//
// A group whose combined radix is RG works on element sets { chunk + row * stride + col :
// row < RG } -- "columns" of the chunk viewed as an RG x stride matrix -- which are closed under
// the group's levels.  A tile is RG rows x CW = 16 (RG = 256: 8) consecutive columns (256 / 128 B segments
// in HBM, so every request is made of full lines), RG CW / 16 threads, 16 c64 per thread; levels inside the
// group exchange through shared memory in natural [row][col] order.  After the last group the
// 256-point base FFTs run as c64_fast_b256_kernel<256,...> on contiguous rows.
//
// Same butterflies and twiddle values as the reference => bit-identical results; the element
// order produced is the reference's (bit-reversed slotting per level).
#include <cstdlib>

#include "c64_dev.cuh"
#include "plan.h"

namespace cfft {
using namespace dev;
namespace {

// One level of radix R whose blocks span SG rows of the tile (RG rows x 16 columns).
//   g/gout : first element of the tile in the HBM source / destination (same buffer when in place),
//            element (row, c) at [row * stride + c]
//   s      : the tile in shared memory, element (row, c) at s[row * CW + c]
//   tw     : planar twiddles of this level, w_k[p] at tw[(k-1) * m + p], m = SG/R * stride
//   col0   : column of the tile's first element inside the chunk
template <int R, int SG, int RG, int CW, bool FWD, bool G_IN, bool G_OUT>
__device__ __forceinline__ void col_level(const c64 *__restrict__ g, c64 *__restrict__ gout, c64 *__restrict__ s,
                                          const c64 *__restrict__ tw, uint32_t stride, uint32_t col0, int t, bool active,
                                          c64 (&v)[16])
{
    constexpr int B = 16 / R, MROW = SG / R;
    const uint32_t m = uint32_t(MROW) * stride;
    int row0[B], col[B];
    uint32_t p[B];
#pragma unroll
    for (int j = 0; j < B; j++) {
        const int b = t + (RG * CW / 16) * j; // butterfly index inside the tile: CW columns x RG/R row-butterflies
        col[j] = b & (CW - 1);
        const int rb = b / CW;
        const int blk = rb / MROW, prow = rb - blk * MROW;
        row0[j] = blk * SG + prow;
        p[j] = uint32_t(prow) * stride + col0 + uint32_t(col[j]);
    }
    if (active || !G_IN) {
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int row = row0[j] + MROW * (FWD ? k : brev_c<R>(k));
                v[j * R + k] = G_IN ? ld_stream(g + size_t(row) * stride + col[j]) : s[row * CW + col[j]];
            }
    }
    if (!G_IN && !G_OUT) __syncthreads(); // in place: all reads before any write
    if (active) {
#pragma unroll
        for (int j = 0; j < B; j++) {
            c64 *x = &v[j * R];
            if (!FWD) {
#pragma unroll
                for (int k = 1; k < R; k++) x[k] = cmul(ld_tw(tw + size_t(k - 1) * m + p[j]), x[k]);
            }
            bfR<R, FWD>(x);
            if (FWD) {
#pragma unroll
                for (int k = 1; k < R; k++) x[k] = cmul(ld_tw(tw + size_t(k - 1) * m + p[j]), x[k]);
            }
        }
    }
    if (active || !G_OUT) {
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) {
                const int row = row0[j] + MROW * (FWD ? brev_c<R>(k) : k);
                if (G_OUT) st_stream(gout + size_t(row) * stride + col[j], v[j * R + k]);
                else s[row * CW + col[j]] = v[j * R + k];
            }
    }
}

// A radix-8 level whose blocks span SG rows followed by the radix-2 level below it (blocks of SG / 8 = 2 rows), fused in
// registers like level_8x2 of c64_fast_kernels.cuh: thread (blk, col) runs BOTH radix-8 butterflies of its block and column
// (prow = 0 and 1), and the radix-2 level pairs exactly output k of the one with output k of the other (rows
// blk SG + 2 brev(k) + {0, 1}), so nothing is exchanged between the two levels and the group (8, 8, 2) of n = 2^15 needs ONE
// trip through shared memory instead of two.  Same butterflies, same twiddles, same order of operations per butterfly
// (src/unordered.rs:24-43, 98-219) => same bits.  Forward: shared memory in, global out; inverse: global in, shared out.
template <int SG, int RG, int CW, bool FWD>
__device__ __forceinline__ void col_level_8x2(const c64 *__restrict__ g, c64 *__restrict__ gout, c64 *__restrict__ s,
                                              const c64 *__restrict__ tw8, const c64 *__restrict__ tw2, uint32_t stride, uint32_t col0,
                                              int t, bool active, c64 (&v)[16])
{
    static_assert(SG == 16, "the radix-2 level below a radix-8 level of 16-row blocks works on 2-row blocks");
    static_assert(RG * CW / 16 == (RG / SG) * CW, "one thread per (block, column)");
    const int col = t & (CW - 1), blk = t / CW;
    const int rbase = blk * SG;
    const uint32_t m8 = 2u * stride;                   // MROW = SG / 8 = 2
    const uint32_t pcol = col0 + uint32_t(col);        // radix-2 twiddle index (prow = 0); radix-8: prow * stride + pcol
    if (FWD) {
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 8; k++) v[j * 8 + k] = s[(rbase + j + 2 * k) * CW + col];
        if (!active) return;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            c64 *x = &v[j * 8];
            bf8<true>(x);
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(tw8 + size_t(k - 1) * m8 + uint32_t(j) * stride + pcol), x[k]);
        }
        const c64 w = ld_tw(tw2 + pcol);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const c64 a = v[k], b = v[8 + k];
            c64 *o = gout + size_t(rbase + 2 * brev_c<8>(k)) * stride + col;
            st_stream(o, cadd(a, b));                    // fwd_butterfly_x2
            st_stream(o + stride, cmul(w, csub(a, b)));
        }
    } else {
        if (active) {
            const c64 w = ld_tw(tw2 + pcol);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const c64 *i = g + size_t(rbase + 2 * brev_c<8>(k)) * stride + col;
                const c64 a = ld_stream(i), b = cmul(w, ld_stream(i + stride)); // inv_butterfly_x2
                v[k] = cadd(a, b);
                v[8 + k] = csub(a, b);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                c64 *x = &v[j * 8];
#pragma unroll
                for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(tw8 + size_t(k - 1) * m8 + uint32_t(j) * stride + pcol), x[k]);
                bf8<false>(x);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 8; k++) s[(rbase + j + 2 * k) * CW + col] = v[j * 8 + k];
    }
}

// all levels of one group on one tile (RG rows x 16 columns); every thread of the CTA must call it
template <int RA, int RB, int RC, int CW, bool FWD>
__device__ __forceinline__ void column_tile(const c64 *__restrict__ g, c64 *__restrict__ go, c64 *__restrict__ s,
                                            const c64 *const (&tw)[3], uint32_t st, uint32_t col0, int t, bool active,
                                            c64 (&v)[16])
{
    constexpr int RG = RA * RB * RC;
    constexpr int SG0 = RG, SG1 = RG / RA, SG2 = RG / (RA * RB);
    if constexpr (RA == 8 && RB == 8 && RC == 2) { // n = 2^15: the last two levels fused in registers, one exchange
        if (FWD) {
            col_level<RA, SG0, RG, CW, true, true, false>(g, go, s, tw[0], st, col0, t, active, v);
            __syncthreads();
            col_level_8x2<SG1, RG, CW, true>(g, go, s, tw[1], tw[2], st, col0, t, active, v);
        } else {
            col_level_8x2<SG1, RG, CW, false>(g, go, s, tw[1], tw[2], st, col0, t, active, v);
            __syncthreads();
            col_level<RA, SG0, RG, CW, false, false, true>(g, go, s, tw[0], st, col0, t, active, v);
        }
        return;
    }
    if (FWD) {
        col_level<RA, SG0, RG, CW, true, true, (RB == 1)>(g, go, s, tw[0], st, col0, t, active, v);
        if (RB > 1) {
            __syncthreads();
            col_level<RB, SG1, RG, CW, true, false, (RC == 1)>(g, go, s, tw[1], st, col0, t, active, v);
        }
        if (RC > 1) {
            __syncthreads();
            col_level<RC, SG2, RG, CW, true, false, true>(g, go, s, tw[2], st, col0, t, active, v);
        }
    } else {
        if (RC > 1) {
            col_level<RC, SG2, RG, CW, false, true, false>(g, go, s, tw[2], st, col0, t, active, v);
            __syncthreads();
            col_level<RB, SG1, RG, CW, false, false, false>(g, go, s, tw[1], st, col0, t, active, v);
            __syncthreads();
            col_level<RA, SG0, RG, CW, false, false, true>(g, go, s, tw[0], st, col0, t, active, v);
        } else if (RB > 1) {
            col_level<RB, SG1, RG, CW, false, true, false>(g, go, s, tw[1], st, col0, t, active, v);
            __syncthreads();
            col_level<RA, SG0, RG, CW, false, false, true>(g, go, s, tw[0], st, col0, t, active, v);
        } else {
            col_level<RA, SG0, RG, CW, false, true, true>(g, go, s, tw[0], st, col0, t, active, v);
        }
    }
}

struct ColParams {
    uint64_t total_tiles;
    uint32_t n;                // transform size
    uint32_t span0;            // span (elements) of the group's first level
    uint32_t stride;           // span0 / RG
    uint32_t tiles_per_chunk;  // stride / CW
    uint32_t tiles_per_row;    // n / (CW RG)
    const c64 *tw[3];          // planar twiddles of the group's levels, outermost first
    uint32_t ahead;            // L2 prefetch distance in tiles (one wave of resident CTAs), 0 = off
};

// Tile width CW: 16 columns (256 B segments) except for RG = 256, where 8 columns (128 B = one full line)
// keep a tile at 128 threads / 32 KiB so that FOUR independent CTAs share an SM instead of two -- the
// same 16 warps, but twice as many independent phases to overlap (n = 2^16: +10-15 %).
template <int RG> struct ColWidth { static constexpr int CW = RG >= 256 ? 8 : 16; };
template <int RG, int CW> struct ColCfg {
    static constexpr int TPT = RG * CW / 16;          // threads per tile (16 c64 per thread)
    static constexpr int NT = TPT < 128 ? 128 : TPT;  // threads per CTA
    static constexpr int TPC = NT / TPT;              // tiles per CTA
    static constexpr int MINB = NT <= 128 ? 4 : 2;
};

template <int RA, int RB, int RC, int CW, bool FWD>
__global__ void __launch_bounds__(ColCfg<RA * RB * RC, CW>::NT, ColCfg<RA * RB * RC, CW>::MINB)
c64_column_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, ColParams prm)
{
    constexpr int RG = RA * RB * RC;
    using Cfg = ColCfg<RG, CW>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lt = threadIdx.x / Cfg::TPT, t = threadIdx.x % Cfg::TPT;
    const uint64_t tile = uint64_t(blockIdx.x) * Cfg::TPC + lt;
    const bool active = tile < prm.total_tiles;
    const uint64_t tl = active ? tile : 0;
    const uint64_t row = tl / prm.tiles_per_row;
    const uint32_t tt = uint32_t(tl - row * prm.tiles_per_row);
    const uint32_t chunk = tt / prm.tiles_per_chunk;
    const uint32_t col0 = (tt - chunk * prm.tiles_per_chunk) * CW;
    const size_t goff = row * prm.n + size_t(chunk) * prm.span0 + col0;
    const c64 *g = src + goff;
    c64 *go = dst + goff;
    c64 *s = reinterpret_cast<c64 *>(smem_raw) + size_t(lt) * RG * CW;
    c64 v[16];
    const uint32_t st = prm.stride;

    // the tile the CTA taking this one's place will start with: RG rows of CW * 16 bytes, requested into L2 now
    if (prm.ahead && tile + prm.ahead < prm.total_tiles) {
        const uint64_t t2 = tile + prm.ahead, row2 = t2 / prm.tiles_per_row;
        const uint32_t tt2 = uint32_t(t2 - row2 * prm.tiles_per_row), chunk2 = tt2 / prm.tiles_per_chunk;
        const c64 *g2 = src + row2 * prm.n + size_t(chunk2) * prm.span0 + (tt2 - chunk2 * prm.tiles_per_chunk) * CW;
        constexpr int LPR = CW / 8; // 128-byte lines per tile row
        for (int i = t; i < RG * LPR; i += Cfg::TPT) asm volatile("prefetch.global.L2 [%0];" ::"l"(g2 + size_t(i / LPR) * st + (i % LPR) * 8));
    }
    column_tile<RA, RB, RC, CW, FWD>(g, go, s, prm.tw, st, col0, t, active, v);
}

template <int RA, int RB, int RC>
cudaError_t launch_group(bool inverse, const c64 *src, c64 *dst, ColParams prm, uint64_t batch, cudaStream_t stream)
{
    constexpr int RG = RA * RB * RC, CW = ColWidth<RG>::CW;
    using Cfg = ColCfg<RG, CW>;
    prm.tiles_per_chunk = prm.stride / CW;
    prm.tiles_per_row = prm.n / (CW * RG);
    prm.total_tiles = batch * prm.tiles_per_row;
    const size_t smem = (RB == 1) ? 0 : size_t(Cfg::NT) * 16 * sizeof(c64);
    static const int env_pf = [] { const char *e = getenv("CFFT_B200_COLUMN_PREFETCH"); return e ? atoi(e) : 1; }();
    prm.ahead = env_pf > 0 ? uint32_t(148 * Cfg::MINB * Cfg::TPC * env_pf) : 0u;
    auto fk = c64_column_kernel<RA, RB, RC, CW, true>;
    auto ik = c64_column_kernel<RA, RB, RC, CW, false>;
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_device != dev) {
            cudaError_t e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(ik, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            configured_device = dev;
        }
    }
    const uint64_t ctas = (prm.total_tiles + Cfg::TPC - 1) / Cfg::TPC;
    if (inverse) ik<<<unsigned(ctas), Cfg::NT, smem, stream>>>(src, dst, prm);
    else fk<<<unsigned(ctas), Cfg::NT, smem, stream>>>(src, dst, prm);
    count_launch();
    return cudaGetLastError();
}


// ---- n = 2^14 .. 2^16: both HBM passes in ONE persistent kernel ---------------------------------
// n = 256 RG is one column group (RG = 64 / 128 / 256 rows x 256 columns) followed by RG base FFTs of
// 256 points.  As two kernels the intermediate makes a round trip through HBM (2 x 2 x 16 x n bytes
// per transform); chunking the batch over several streams keeps part of it in the 126 MB L2 but the
// schedule becomes launch-bound (hundreds of 4 us kernels per call).  Here resident CTAs pull work items
// from one queue in which the second-phase items of transform j sit `lag` transforms behind its
// first-phase items: by the time a CTA picks them up the first phase of j has normally finished
// (a per-transform counter, release / acquire, makes that a guarantee) and its output -- a few MiB
// back in the write stream -- is still in L2.  HBM then sees each element once in and once out.
// Deadlock-free: items are handed out in queue order to CTAs that are already running, and an item
// only ever waits for items handed out before it.
struct TwoPassParams {
    c64 *data;
    uint32_t batch, n, lag;
    const c64 *tw[3];   // planar twiddles of the column group's levels, outermost first
    const c64 *tw_base; // planar half of init_wt(16, 256)
    uint32_t *sync;     // [0] queue head; [1 + j] finished first-phase items of transform j
    uint32_t prefetch;  // L2 prefetch of the first-phase item one wave ahead (CFFT_B200_TWOPASS_PREFETCH, default on)
};

__device__ unsigned int g_twopass_timeouts = 0; // second-phase items that gave up waiting for their first phase

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int RA, int RB, int RC, bool FWD>
__global__ void __launch_bounds__(ColCfg<RA * RB * RC, ColWidth<RA * RB * RC>::CW>::NT, ColCfg<RA * RB * RC, ColWidth<RA * RB * RC>::CW>::MINB)
c64_twopass_kernel(TwoPassParams prm)
{
    constexpr int RG = RA * RB * RC, CW = ColWidth<RG>::CW;
    using Cfg = ColCfg<RG, CW>;
    constexpr int NT = Cfg::NT;
    constexpr int COL_ITEMS = (256 / CW) / Cfg::TPC;         // 256 columns per transform in tiles of CW
    constexpr int ROWS_PER_ITEM = NT / 16, ROW_ITEMS = RG / ROWS_PER_ITEM;
    constexpr int F_ITEMS = FWD ? COL_ITEMS : ROW_ITEMS;     // forward: columns first; inverse: rows first
    constexpr int W = COL_ITEMS + ROW_ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t sh_q[2];
    c64 *smem = reinterpret_cast<c64 *>(smem_raw);
    const uint32_t total_q = (prm.batch + prm.lag) * W;
    c64 v[16];

    // Thread 0 keeps one queue ticket in flight: the ticket for the NEXT item is requested when an item
    // starts and published (shared memory) when it ends, so the atomic's round trip to L2 is off the
    // critical path.  One block barrier per item: it publishes the ticket, frees the shared-memory tile
    // and orders every thread's stores before thread 0 signals the finished first-phase item.
    uint32_t next_q = 0, signal_j = 0;
    bool signal = false;
    int buf = 0;
    if (threadIdx.x == 0) sh_q[0] = atomicAdd(prm.sync, 1u);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0 && signal) {
            __threadfence();
            atomicAdd(prm.sync + 1 + signal_j, 1u);
            signal = false;
        }
        const uint32_t q = sh_q[buf];
        if (q >= total_q) break;
        if (threadIdx.x == 0) next_q = atomicAdd(prm.sync, 1u);
        const uint32_t step = q / W, r = q - step * W;
        const bool first = r < F_ITEMS;
        const bool valid = first ? step < prm.batch : step >= prm.lag;
        if (prm.prefetch) {
            // The item one wave of CTAs further down the queue will be picked up by SOME CTA about one item time from now.
            // If it is a first-phase item its data still sits in HBM: ask L2 for it now (second-phase items read what the
            // first phase has just written, i.e. L2 already).  A hint only: nothing depends on who ends up with the item.
            const uint32_t q2 = q + gridDim.x, step2 = q2 / W, r2 = q2 - step2 * W;
            if (r2 < uint32_t(F_ITEMS) && step2 < prm.batch) {
                const c64 *row2 = prm.data + size_t(step2) * prm.n;
                if (FWD) { // column item: TPC tiles of RG rows x CW columns, rows 256 elements apart
                    constexpr int LPR = CW / 8, LINES = Cfg::TPC * RG * LPR; // 128-byte lines per row / per item
                    for (int i = threadIdx.x; i < LINES; i += NT) {
                        const int tile = i / (RG * LPR), rem = i - tile * (RG * LPR), rr = rem / LPR, l = rem - rr * LPR;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row2 + (r2 * Cfg::TPC + tile) * CW + size_t(rr) * 256 + l * 8));
                    }
                } else { // rows item: ROWS_PER_ITEM * 256 contiguous elements
                    for (int i = threadIdx.x; i < ROWS_PER_ITEM * 32; i += NT)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(row2 + size_t(r2) * ROWS_PER_ITEM * 256 + i * 8));
                }
            }
        }
        if (valid) {
            const uint32_t j = first ? step : step - prm.lag;
            const uint32_t sub = first ? r : r - F_ITEMS;
            if (!first) {
                if (threadIdx.x == 0) {
                    uint32_t spins = 0;
                    while (ld_acquire_u32(prm.sync + 1 + j) < uint32_t(F_ITEMS)) {
                        __nanosleep(200);
                        // ~7 s without the signal (producers descheduled by MPS / a debugger / preemption): never hang and
                        // never trap the context -- give up waiting, count it, and let the host see the count
                        // (cfft_twopass_timeouts); the variant is opt-in for exactly this reason
                        if (++spins > (1u << 25)) {
                            atomicAdd(&g_twopass_timeouts, 1u);
                            break;
                        }
                    }
                }
                __syncthreads();
            }
            c64 *row = prm.data + size_t(j) * prm.n;
            if (first == FWD) { // column item: Cfg::TPC tiles
                const int lt = threadIdx.x / Cfg::TPT, t = threadIdx.x % Cfg::TPT;
                const uint32_t col0 = (sub * Cfg::TPC + lt) * CW;
                column_tile<RA, RB, RC, CW, FWD>(row + col0, row + col0, smem + size_t(lt) * RG * CW, prm.tw, 256u, col0, t, true, v);
            } else { // rows item: ROWS_PER_ITEM base FFTs
                const int hw = threadIdx.x / 16, lane16 = threadIdx.x % 16;
                c64 *gp = row + size_t(sub * ROWS_PER_ITEM + hw) * 256;
                base256<FWD, true, true>(gp, smem + hw * 256, gp, prm.tw_base, lane16, v);
            }
            signal = first;
            signal_j = j;
        }
        buf ^= 1;
        if (threadIdx.x == 0) sh_q[buf] = next_q;
    }
}

template <int RA, int RB, int RC>
cudaError_t launch_twopass(bool inverse, const TwoPassParams &prm_in, int device, cudaStream_t stream)
{
    constexpr int RG = RA * RB * RC, CW = ColWidth<RG>::CW;
    using Cfg = ColCfg<RG, CW>;
    constexpr int W = (256 / CW) / Cfg::TPC + RG / (Cfg::NT / 16);
    const size_t smem = size_t(Cfg::NT) * 16 * sizeof(c64);
    auto fk = c64_twopass_kernel<RA, RB, RC, true>;
    auto ik = c64_twopass_kernel<RA, RB, RC, false>;
    static thread_local int configured_device = -1;
    static thread_local int resident = 0;
    if (configured_device != device) {
        cudaError_t e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ik, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        int sms = 0, occ_f = 0, occ_i = 0;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, fk, Cfg::NT, smem);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, ik, Cfg::NT, smem);
        if (e != cudaSuccess) return e;
        resident = sms * (occ_f < occ_i ? occ_f : occ_i);
        if (resident < 1) return cudaErrorLaunchOutOfResources;
        configured_device = device;
    }
    TwoPassParams prm = prm_in;
    static const int env_pf = [] { const char *e = getenv("CFFT_B200_TWOPASS_PREFETCH"); return e ? atoi(e) : 1; }();
    prm.prefetch = (env_pf > 0 && !inverse) ? 1u : 0u; // measured: forward +1 .. +3 %, inverse -1 .. -3 % (profiles/r2e_prefetch_ab.txt)
    if (prm.lag == 0) prm.lag = uint32_t((3 * resident / 2 + W - 1) / W); // ~1.5 waves of items between the phases
    if (prm.lag > prm.batch) prm.lag = prm.batch;
    if (uint64_t(prm.batch) + prm.lag > (0xFFFFFFFFull / W)) return cudaErrorInvalidValue;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(device, &pool);
    if (e != cudaSuccess) return e;
    const size_t sync_bytes = (size_t(prm.batch) + 1) * sizeof(uint32_t);
    if ((e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&prm.sync), sync_bytes, pool, stream)) != cudaSuccess) return e;
    e = cudaMemsetAsync(prm.sync, 0, sync_bytes, stream);
    if (e == cudaSuccess) {
        const uint64_t items = (uint64_t(prm.batch) + prm.lag) * W;
        const unsigned ctas = unsigned(items < uint64_t(resident) ? items : uint64_t(resident));
        if (inverse) ik<<<ctas, Cfg::NT, smem, stream>>>(prm);
        else fk<<<ctas, Cfg::NT, smem, stream>>>(prm);
        count_launch();
        e = cudaGetLastError();
    }
    const cudaError_t e2 = cudaFreeAsync(prm.sync, stream);
    return e != cudaSuccess ? e : e2;
}

} // namespace

// radices: outermost level first; 1 = absent.  src == dst: in place; else out of place (ordered plans)
cudaError_t launch_c64_column_group(bool inverse, const double2 *src, double2 *dst, uint64_t batch, uint32_t n, uint32_t span0,
                                    const int radices[3], const double2 *const tw[3], cudaStream_t stream)
{
    const int ra = radices[0], rb = radices[1], rc = radices[2];
    const uint32_t rg = uint32_t(ra * rb * rc);
    ColParams prm;
    prm.n = n;
    prm.span0 = span0;
    prm.stride = span0 / rg;
    prm.tiles_per_chunk = prm.tiles_per_row = 0; // filled in per tile width by launch_group
    prm.total_tiles = 0;
    prm.ahead = 0;
    for (int i = 0; i < 3; i++) prm.tw[i] = tw[i];
    const int key = ra * 100 + rb * 10 + rc;
    // two radix-8 levels: one thread per column with tensor memory between the levels (c64_tmem.cu), bit-identical.
    // Measured slower than the shared-memory tiles on B200 (profiles/r2a_tmem_columns.txt: n = 2^16 3.0 vs 3.3 TB/s:
    // eight warps per SM with one serial chain each do not keep HBM busy), so it is opt-in: CFFT_B200_TMEM_COLUMNS=1
    static const bool use_tmem = [] { const char *e = getenv("CFFT_B200_TMEM_COLUMNS"); return e && atoi(e) != 0; }();
    if (key == 881 && use_tmem && tmem_column88_supported(n, span0))
        return launch_c64_tmem_column88(inverse, src, dst, batch, n, span0, tw[0], tw[1], stream);
    // groups of two or three levels: optionally the persistent kernel of c64_colpipe.cu (CFFT_B200_COLPIPE_MIN_BATCH:
    // smallest batch per launch that takes it; same bits either way)
    // measured on the B200 (profiles/r2c_variants*.txt): 3-8 % SLOWER than the one-shot tiles in every schedule (244 registers
    // -> eight warps per SM; persistent grids sized for the whole machine also serialise the multi-stream chunked
    // schedule), so it is opt-in: CFFT_B200_COLPIPE=1
    static const bool use_pipe = [] { const char *e = getenv("CFFT_B200_COLPIPE"); return e && atoi(e) != 0; }();
    static const long pipe_min_batch = [] { const char *e = getenv("CFFT_B200_COLPIPE_MIN_BATCH"); return e ? atol(e) : 1; }();
    if (use_pipe && colpipe_supported(radices) && long(batch) >= pipe_min_batch) {
        int dev = 0;
        cudaError_t ce = cudaGetDevice(&dev);
        if (ce != cudaSuccess) return ce;
        return launch_c64_colpipe_group(inverse, src, dst, batch, n, span0, radices, tw, dev, stream);
    }
    switch (key) {
    case 811: return launch_group<8, 1, 1>(inverse, src, dst, prm, batch, stream);
    case 411: return launch_group<4, 1, 1>(inverse, src, dst, prm, batch, stream);
    case 211: return launch_group<2, 1, 1>(inverse, src, dst, prm, batch, stream);
    case 821: return launch_group<8, 2, 1>(inverse, src, dst, prm, batch, stream);
    case 841: return launch_group<8, 4, 1>(inverse, src, dst, prm, batch, stream);
    case 881: return launch_group<8, 8, 1>(inverse, src, dst, prm, batch, stream);
    case 882: return launch_group<8, 8, 2>(inverse, src, dst, prm, batch, stream);
    case 884: return launch_group<8, 8, 4>(inverse, src, dst, prm, batch, stream);
    case 888: return launch_group<8, 8, 8>(inverse, src, dst, prm, batch, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cfft

namespace cfft {

// number of second-phase work items of the persistent kernel that timed out waiting on this device since load (0 in a
// healthy process; a non-zero count means results of those calls are invalid and variant 2 should be used instead)
cudaError_t twopass_timeouts(unsigned int *out)
{
    return cudaMemcpyFromSymbol(out, g_twopass_timeouts, sizeof(unsigned int));
}

// n = 2^14, 2^15, 2^16 (one column group + base FFTs): the persistent two-phase kernel.  lag = 0: automatic.
cudaError_t launch_c64_twopass(bool inverse, double2 *data, uint64_t batch, uint32_t n, const int radices[3],
                               const double2 *const tw[3], const double2 *tw_base, uint32_t lag, int device, cudaStream_t stream)
{
    if (batch >= (uint64_t{1} << 26)) return cudaErrorInvalidValue;
    TwoPassParams prm;
    prm.data = data;
    prm.batch = uint32_t(batch);
    prm.n = n;
    prm.lag = lag;
    for (int i = 0; i < 3; i++) prm.tw[i] = tw[i];
    prm.tw_base = tw_base;
    prm.sync = nullptr;
    prm.prefetch = 0;
    const int key = radices[0] * 100 + radices[1] * 10 + radices[2];
    if (n == 16384 && key == 881) return launch_twopass<8, 8, 1>(inverse, prm, device, stream);
    if (n == 32768 && key == 882) return launch_twopass<8, 8, 2>(inverse, prm, device, stream);
    if (n == 65536 && key == 884) return launch_twopass<8, 8, 4>(inverse, prm, device, stream);
    return cudaErrorInvalidValue;
}

} // namespace cfft

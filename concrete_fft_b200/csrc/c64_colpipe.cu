// c64_colpipe.cu -- the unordered levels of large transforms (n >= 2^14), PERSISTENT form of c64_column.cu.
//
// Same tiles, same thread -> butterfly maps, same butterflies and twiddle values as c64_column_kernel (so the same bits,
// fwd_process_x* / inv_process_x*, src/unordered.rs:222-293), but a CTA no longer handles one tile and exits.  ncu on the
// one-shot kernel with its data L2-resident (profiles/r2b_l2res_*.txt) shows no saturated unit -- L1 60-70 %, FP64 27-30 %,
// issue slots 30 % -- and the waiting spread over: the tile's own loads (17 %), the level twiddles arriving from L2 (9 %:
// the outer level's table is n * 14 bytes, far beyond L1), stores draining before the CTA may exit (14 %), index set-up.
// Here a CTA keeps ONE tile position and walks through the batch:
//   * the twiddles of a position never change, so they are loaded once and stay in registers (outer level: 14 c64 per
//     thread; both levels of a two-level group): no twiddle traffic and no twiddle latency inside the loop;
//   * the next transform's 16 elements per thread are requested (128-bit streaming loads into a second register set)
//     before the current tile is computed, so a full tile of loads is always in flight per thread;
//   * the exchange buffer is double-buffered: levels - 1 barriers per tile, between the 64 .. 256 threads of ONE tile only
//     (named barriers), and stores are never waited for.
// 212-255 registers -> two CTAs of 128 threads (or one of 256) per SM: eight warps, each with a tile's worth of
// independent work.
#include <cstdlib>

#include "c64_dev.cuh"
#include "plan.h"

namespace cfft {
using namespace dev;
namespace {

template <int R, int SG, int RG, int CW> struct LevelMap {
    static constexpr int B = 16 / R, MROW = SG / R, TPT = RG * CW / 16;
    // butterfly j of thread t covers rows row0 + MROW * k (k < R) of column col; its twiddles are w_k[prow * stride + col0 + col]
    __device__ __forceinline__ static void get(int t, int j, int &row0, int &col, int &prow)
    {
        const int b = t + TPT * j;
        col = b & (CW - 1);
        const int rb = b / CW;
        const int blk = rb / MROW;
        prow = rb - blk * MROW;
        row0 = blk * SG + prow;
    }
};

enum { IO_REGS = 0, IO_SMEM = 1, IO_GLOBAL = 2 };

// One level on the 16 values of a thread.  IN: IO_REGS (v already holds the level's inputs, loaded by prefetch_tile) or
// IO_SMEM; OUT: IO_SMEM or IO_GLOBAL.  TWR: twiddles in registers (twr[j][k-1]) or from the planar table (tw, L1-resident
// for the inner levels of a fixed position).
template <int R, int SG, int RG, int CW, bool FWD, int IN, int OUT, bool TWR>
__device__ __forceinline__ void plevel(const c64 *__restrict__ sin, c64 *__restrict__ sout, c64 *__restrict__ gout, size_t stride,
                                       const c64 *__restrict__ tw, uint32_t col0, const c64 (&twr)[16 / R][R - 1], int t, c64 (&v)[16])
{
    using M = LevelMap<R, SG, RG, CW>;
    constexpr int B = M::B, MROW = M::MROW;
    int row0[B], col[B], prow[B];
#pragma unroll
    for (int j = 0; j < B; j++) M::get(t, j, row0[j], col[j], prow[j]);
    if (IN == IO_SMEM) {
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) v[j * R + k] = sin[(row0[j] + MROW * (FWD ? k : brev_c<R>(k))) * CW + col[j]];
    }
#pragma unroll
    for (int j = 0; j < B; j++) {
        c64 *x = &v[j * R];
        c64 w[R - 1];
#pragma unroll
        for (int k = 1; k < R; k++)
            w[k - 1] = TWR ? twr[j][k - 1] : ld_tw(tw + size_t(k - 1) * (size_t(MROW) * stride) + size_t(prow[j]) * stride + col0 + col[j]);
        if (!FWD) {
#pragma unroll
            for (int k = 1; k < R; k++) x[k] = cmul(w[k - 1], x[k]);
        }
        bfR<R, FWD>(x);
        if (FWD) {
#pragma unroll
            for (int k = 1; k < R; k++) x[k] = cmul(w[k - 1], x[k]);
        }
    }
#pragma unroll
    for (int j = 0; j < B; j++)
#pragma unroll
        for (int k = 0; k < R; k++) {
            const int row = row0[j] + MROW * (FWD ? brev_c<R>(k) : k);
            if (OUT == IO_GLOBAL) st_stream(gout + size_t(row) * stride + col[j], v[j * R + k]);
            else sout[row * CW + col[j]] = v[j * R + k];
        }
}

// the 16 inputs of the FIRST executed level of a tile, straight from HBM / L2 into registers
template <int R, int SG, int RG, int CW, bool FWD>
__device__ __forceinline__ void prefetch_tile(const c64 *__restrict__ g, size_t stride, int t, c64 (&v)[16])
{
    using M = LevelMap<R, SG, RG, CW>;
#pragma unroll
    for (int j = 0; j < M::B; j++) {
        int row0, col, prow;
        M::get(t, j, row0, col, prow);
#pragma unroll
        for (int k = 0; k < R; k++) v[j * R + k] = ld_stream(g + size_t(row0 + M::MROW * (FWD ? k : brev_c<R>(k))) * stride + col);
    }
}

template <int R, int SG, int RG, int CW>
__device__ __forceinline__ void load_twiddles(const c64 *__restrict__ tw, size_t stride, uint32_t col0, int t, c64 (&twr)[16 / R][R - 1])
{
    using M = LevelMap<R, SG, RG, CW>;
#pragma unroll
    for (int j = 0; j < M::B; j++) {
        int row0, col, prow;
        M::get(t, j, row0, col, prow);
#pragma unroll
        for (int k = 1; k < R; k++) twr[j][k - 1] = ld_tw(tw + size_t(k - 1) * (size_t(M::MROW) * stride) + size_t(prow) * stride + col0 + col);
    }
}

struct PipeParams {
    uint32_t batch;
    uint32_t n;               // transform size
    uint32_t span0;           // span of the group's first level
    uint32_t stride;          // span0 / RG
    uint32_t tiles_per_chunk; // stride / CW
    uint32_t positions;       // CTA positions per transform = n / (CW RG TPC)
    uint32_t per_position;    // CTAs sharing one position (they take transforms j = seq, seq + per_position, ...)
    const c64 *tw[3];
};

template <int RG> struct PipeWidth { static constexpr int CW = RG >= 256 ? 8 : 16; };
template <int RG, int CW> struct PipeCfg {
    static constexpr int TPT = RG * CW / 16;         // threads per tile
    static constexpr int NT = TPT < 128 ? 128 : TPT; // threads per CTA
    static constexpr int TPC = NT / TPT;             // tiles per CTA, side by side (TPC * CW consecutive columns)
    static constexpr int MINB = NT <= 128 ? 2 : 1;
    static constexpr size_t SMEM = size_t(2) * TPC * RG * CW * sizeof(c64); // two exchange buffers per tile
};

template <int RA, int RB, int RC, int CW, bool FWD>
__global__ void __launch_bounds__(PipeCfg<RA * RB * RC, CW>::NT, PipeCfg<RA * RB * RC, CW>::MINB)
c64_colpipe_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, PipeParams prm)
{
    constexpr int RG = RA * RB * RC;
    using Cfg = PipeCfg<RG, CW>;
    constexpr int SG0 = RG, SG1 = RG / RA, SG2 = RG / (RA * RB);
    constexpr int NLEV = RC > 1 ? 3 : (RB > 1 ? 2 : 1);
    static_assert(NLEV >= 2, "single-level groups have nothing to exchange: c64_column_kernel");
    constexpr bool TWB_REGS = NLEV == 2; // two-level groups keep both levels' twiddles in registers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lt = threadIdx.x / Cfg::TPT, t = threadIdx.x % Cfg::TPT;
    const uint32_t pos = blockIdx.x % prm.positions, seq = blockIdx.x / prm.positions;
    const uint32_t tt = pos * Cfg::TPC + uint32_t(lt);
    const uint32_t chunk = tt / prm.tiles_per_chunk;
    const uint32_t col0 = (tt - chunk * prm.tiles_per_chunk) * CW;
    const size_t goff = size_t(chunk) * prm.span0 + col0;
    const size_t st = prm.stride;
    c64 *s0 = reinterpret_cast<c64 *>(smem_raw) + size_t(lt) * 2 * RG * CW;
    c64 *s1 = s0 + RG * CW;
    auto tile_barrier = [&] {
        if (Cfg::TPC == 1) __syncthreads();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + lt), "r"(Cfg::TPT) : "memory");
    };

    constexpr int RCX = RC == 1 ? 2 : RC; // a type for the third level's (unused) register twiddles when the group has two levels
    c64 twa[16 / RA][RA - 1];
    c64 twb[16 / RB][RB - 1];
    c64 twc[16 / RCX][RCX - 1]; // never loaded: the innermost of three levels reads its few twiddles from L1
    load_twiddles<RA, SG0, RG, CW>(prm.tw[0], st, col0, t, twa);
    if constexpr (TWB_REGS) load_twiddles<RB, SG1, RG, CW>(prm.tw[1], st, col0, t, twb);

    auto prefetch = [&](uint32_t jj, c64(&dstv)[16]) {
        const c64 *g = src + size_t(jj) * prm.n + goff;
        if constexpr (FWD) prefetch_tile<RA, SG0, RG, CW, true>(g, st, t, dstv);
        else if constexpr (NLEV == 3) prefetch_tile<RCX, SG2, RG, CW, false>(g, st, t, dstv);
        else prefetch_tile<RB, SG1, RG, CW, false>(g, st, t, dstv);
    };

    c64 v[16], nxt[16];
    uint32_t j = seq;
    if (j < prm.batch) prefetch(j, nxt);
    for (; j < prm.batch; j += prm.per_position) {
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = nxt[i];
        const uint32_t jn = j + prm.per_position;
        if (jn < prm.batch) prefetch(jn, nxt); // the next tile of this position is on its way while this one is computed
        c64 *go = dst + size_t(j) * prm.n + goff;
        if constexpr (NLEV == 2) { // one exchange per tile: alternate the two buffers between tiles (a thread may start
            c64 *tmp = s0;         // writing tile i + 1 while another still reads tile i)
            s0 = s1;
            s1 = tmp;
        }
        if constexpr (FWD) {
            plevel<RA, SG0, RG, CW, true, IO_REGS, IO_SMEM, true>(nullptr, s0, nullptr, st, prm.tw[0], col0, twa, t, v);
            tile_barrier();
            if constexpr (NLEV == 2) {
                plevel<RB, SG1, RG, CW, true, IO_SMEM, IO_GLOBAL, TWB_REGS>(s0, nullptr, go, st, prm.tw[1], col0, twb, t, v);
            } else {
                plevel<RB, SG1, RG, CW, true, IO_SMEM, IO_SMEM, TWB_REGS>(s0, s1, nullptr, st, prm.tw[1], col0, twb, t, v);
                tile_barrier();
                plevel<RCX, SG2, RG, CW, true, IO_SMEM, IO_GLOBAL, false>(s1, nullptr, go, st, prm.tw[2], col0, twc, t, v);
            }
        } else {
            if constexpr (NLEV == 2) {
                plevel<RB, SG1, RG, CW, false, IO_REGS, IO_SMEM, TWB_REGS>(nullptr, s0, nullptr, st, prm.tw[1], col0, twb, t, v);
                tile_barrier();
                plevel<RA, SG0, RG, CW, false, IO_SMEM, IO_GLOBAL, true>(s0, nullptr, go, st, prm.tw[0], col0, twa, t, v);
            } else {
                plevel<RCX, SG2, RG, CW, false, IO_REGS, IO_SMEM, false>(nullptr, s0, nullptr, st, prm.tw[2], col0, twc, t, v);
                tile_barrier();
                plevel<RB, SG1, RG, CW, false, IO_SMEM, IO_SMEM, TWB_REGS>(s0, s1, nullptr, st, prm.tw[1], col0, twb, t, v);
                tile_barrier();
                plevel<RA, SG0, RG, CW, false, IO_SMEM, IO_GLOBAL, true>(s1, nullptr, go, st, prm.tw[0], col0, twa, t, v);
            }
        }
    }
}

template <int RA, int RB, int RC>
cudaError_t launch_pipe(bool inverse, const c64 *src, c64 *dst, PipeParams prm, int device, cudaStream_t stream)
{
    constexpr int RG = RA * RB * RC, CW = PipeWidth<RG>::CW;
    using Cfg = PipeCfg<RG, CW>;
    prm.tiles_per_chunk = prm.stride / CW;
    prm.positions = prm.n / (CW * RG * Cfg::TPC);
    auto fk = c64_colpipe_kernel<RA, RB, RC, CW, true>;
    auto ik = c64_colpipe_kernel<RA, RB, RC, CW, false>;
    static thread_local int configured_device = -1;
    static thread_local int resident = 0;
    if (configured_device != device) {
        cudaError_t e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ik, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::SMEM));
        int sms = 0, occ_f = 0, occ_i = 0;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, fk, Cfg::NT, Cfg::SMEM);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_i, ik, Cfg::NT, Cfg::SMEM);
        if (e != cudaSuccess) return e;
        resident = sms * (occ_f < occ_i ? occ_f : occ_i);
        if (resident < 1) return cudaErrorLaunchOutOfResources;
        configured_device = device;
    }
    // CTAs per position: fill the machine once (all CTAs resident), never more CTAs per position than transforms
    uint32_t per = uint32_t(resident) / prm.positions;
    if (per < 1) per = 1;
    if (per > prm.batch) per = prm.batch;
    prm.per_position = per;
    const unsigned grid = prm.positions * per;
    if (inverse) ik<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(src, dst, prm);
    else fk<<<grid, Cfg::NT, Cfg::SMEM, stream>>>(src, dst, prm);
    count_launch();
    return cudaGetLastError();
}

} // namespace

bool colpipe_supported(const int radices[3])
{
    const int key = radices[0] * 100 + radices[1] * 10 + radices[2];
    return key == 881 || key == 882 || key == 884 || key == 888 || key == 841;
}

// radices: outermost level first; 1 = absent.  src == dst: in place; else out of place (ordered plans)
cudaError_t launch_c64_colpipe_group(bool inverse, const double2 *src, double2 *dst, uint64_t batch, uint32_t n, uint32_t span0,
                                     const int radices[3], const double2 *const tw[3], int device, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    if (batch > 0xFFFFFFFFull) return cudaErrorInvalidValue;
    const int ra = radices[0], rb = radices[1], rc = radices[2];
    PipeParams prm;
    prm.batch = uint32_t(batch);
    prm.n = n;
    prm.span0 = span0;
    prm.stride = span0 / uint32_t(ra * rb * rc);
    prm.tiles_per_chunk = prm.positions = prm.per_position = 0;
    for (int i = 0; i < 3; i++) prm.tw[i] = tw[i];
    switch (ra * 100 + rb * 10 + rc) {
    case 841: return launch_pipe<8, 4, 1>(inverse, src, dst, prm, device, stream);
    case 881: return launch_pipe<8, 8, 1>(inverse, src, dst, prm, device, stream);
    case 882: return launch_pipe<8, 8, 2>(inverse, src, dst, prm, device, stream);
    case 884: return launch_pipe<8, 8, 4>(inverse, src, dst, prm, device, stream);
    case 888: return launch_pipe<8, 8, 8>(inverse, src, dst, prm, device, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cfft

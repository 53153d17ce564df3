// c64_column.cu -- the unordered levels of large transforms (n >= 2^14) for plans with base
// (Dif16, 256): groups of up to three consecutive radix-2/4/8 levels
// (fwd_process_x* / inv_process_x*, src/unordered.rs:222-293) executed in one HBM pass.
//
// A group whose combined radix is RG works on element sets { chunk + row * stride + col :
// row < RG } -- "columns" of the chunk viewed as an RG x stride matrix -- which are closed under
// the group's levels.  A tile is RG rows x 16 consecutive columns (256 B segments in HBM, so
// every request is made of full sectors), RG threads, 16 c64 per thread; levels inside the
// group exchange through shared memory in natural [row][col] order.  After the last group the
// 256-point base FFTs run as c64_fast_b256_kernel<256,...> on contiguous rows.
//
// Same butterflies and twiddle values as the reference => bit-identical results; the element
// order produced is the reference's (bit-reversed slotting per level).
#include "c64_math.cuh"
#include "plan.h"

namespace cfft {
namespace {

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

template <int R> __device__ __forceinline__ constexpr int brev_c(int k)
{
    return R == 2 ? k : (R == 4 ? ((k & 1) << 1) | (k >> 1) : ((k & 1) << 2) | (k & 2) | (k >> 2));
}

// One level of radix R whose blocks span SG rows of the tile (RG rows x 16 columns).
//   g/gout : first element of the tile in the HBM source / destination (same buffer when in place),
//            element (row, c) at [row * stride + c]
//   s      : the tile in shared memory, element (row, c) at s[row * 16 + c]
//   tw     : planar twiddles of this level, w_k[p] at tw[(k-1) * m + p], m = SG/R * stride
//   col0   : column of the tile's first element inside the chunk
template <int R, int SG, int RG, bool FWD, bool G_IN, bool G_OUT>
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
        const int b = t + RG * j;  // butterfly index inside the tile: 16 columns x RG/R row-butterflies
        col[j] = b & 15;
        const int rb = b >> 4;
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
                v[j * R + k] = G_IN ? ld_stream(g + size_t(row) * stride + col[j]) : s[row * 16 + col[j]];
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
                else s[row * 16 + col[j]] = v[j * R + k];
            }
    }
}

struct ColParams {
    uint64_t total_tiles;
    uint32_t n;                // transform size
    uint32_t span0;            // span (elements) of the group's first level
    uint32_t stride;           // span0 / RG
    uint32_t tiles_per_chunk;  // stride / 16
    uint32_t tiles_per_row;    // n / (16 RG)
    const c64 *tw[3];          // planar twiddles of the group's levels, outermost first
};

template <int RG> struct ColCfg {
    static constexpr int NT = RG < 128 ? 128 : RG; // threads per CTA
    static constexpr int TPC = NT / RG;            // tiles per CTA
    static constexpr int MINB = NT <= 128 ? 4 : 2;
};

template <int RA, int RB, int RC, bool FWD>
__global__ void __launch_bounds__(ColCfg<RA * RB * RC>::NT, ColCfg<RA * RB * RC>::MINB)
c64_column_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, ColParams prm)
{
    constexpr int RG = RA * RB * RC;
    using Cfg = ColCfg<RG>;
    constexpr int SG0 = RG, SG1 = RG / RA, SG2 = RG / (RA * RB);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lt = threadIdx.x / RG, t = threadIdx.x % RG;
    const uint64_t tile = uint64_t(blockIdx.x) * Cfg::TPC + lt;
    const bool active = tile < prm.total_tiles;
    const uint64_t tl = active ? tile : 0;
    const uint64_t row = tl / prm.tiles_per_row;
    const uint32_t tt = uint32_t(tl - row * prm.tiles_per_row);
    const uint32_t chunk = tt / prm.tiles_per_chunk;
    const uint32_t col0 = (tt - chunk * prm.tiles_per_chunk) * 16;
    const size_t goff = row * prm.n + size_t(chunk) * prm.span0 + col0;
    const c64 *g = src + goff;
    c64 *go = dst + goff;
    c64 *s = reinterpret_cast<c64 *>(smem_raw) + size_t(lt) * RG * 16;
    c64 v[16];
    const uint32_t st = prm.stride;

    if (FWD) {
        col_level<RA, SG0, RG, true, true, (RB == 1)>(g, go, s, prm.tw[0], st, col0, t, active, v);
        if (RB > 1) {
            __syncthreads();
            col_level<RB, SG1, RG, true, false, (RC == 1)>(g, go, s, prm.tw[1], st, col0, t, active, v);
        }
        if (RC > 1) {
            __syncthreads();
            col_level<RC, SG2, RG, true, false, true>(g, go, s, prm.tw[2], st, col0, t, active, v);
        }
    } else {
        if (RC > 1) {
            col_level<RC, SG2, RG, false, true, false>(g, go, s, prm.tw[2], st, col0, t, active, v);
            __syncthreads();
            col_level<RB, SG1, RG, false, false, false>(g, go, s, prm.tw[1], st, col0, t, active, v);
            __syncthreads();
            col_level<RA, SG0, RG, false, false, true>(g, go, s, prm.tw[0], st, col0, t, active, v);
        } else if (RB > 1) {
            col_level<RB, SG1, RG, false, true, false>(g, go, s, prm.tw[1], st, col0, t, active, v);
            __syncthreads();
            col_level<RA, SG0, RG, false, false, true>(g, go, s, prm.tw[0], st, col0, t, active, v);
        } else {
            col_level<RA, SG0, RG, false, true, true>(g, go, s, prm.tw[0], st, col0, t, active, v);
        }
    }
}

template <int RA, int RB, int RC>
cudaError_t launch_group(bool inverse, const c64 *src, c64 *dst, const ColParams &prm, cudaStream_t stream)
{
    constexpr int RG = RA * RB * RC;
    using Cfg = ColCfg<RG>;
    const size_t smem = (RB == 1) ? 0 : size_t(Cfg::NT) * 16 * sizeof(c64);
    auto fk = c64_column_kernel<RA, RB, RC, true>;
    auto ik = c64_column_kernel<RA, RB, RC, false>;
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
    prm.tiles_per_chunk = prm.stride / 16;
    prm.tiles_per_row = n / (16 * rg);
    prm.total_tiles = uint64_t(batch) * prm.tiles_per_row;
    for (int i = 0; i < 3; i++) prm.tw[i] = tw[i];
    const int key = ra * 100 + rb * 10 + rc;
    switch (key) {
    case 811: return launch_group<8, 1, 1>(inverse, src, dst, prm, stream);
    case 411: return launch_group<4, 1, 1>(inverse, src, dst, prm, stream);
    case 211: return launch_group<2, 1, 1>(inverse, src, dst, prm, stream);
    case 821: return launch_group<8, 2, 1>(inverse, src, dst, prm, stream);
    case 841: return launch_group<8, 4, 1>(inverse, src, dst, prm, stream);
    case 881: return launch_group<8, 8, 1>(inverse, src, dst, prm, stream);
    case 882: return launch_group<8, 8, 2>(inverse, src, dst, prm, stream);
    case 884: return launch_group<8, 8, 4>(inverse, src, dst, prm, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cfft

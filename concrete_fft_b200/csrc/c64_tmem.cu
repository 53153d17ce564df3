// c64_tmem.cu -- unordered levels with TENSOR MEMORY as the parking space between two levels.
//
// Two consecutive radix-8 levels (fwd_process_x8 / inv_process_x8, src/unordered.rs:266-293) of a large
// transform work on "columns": 64 elements { chunk + row * stride + col : row < 64 } that are closed under both
// levels.  c64_column.cu gives 16 elements to a thread and exchanges between the levels through shared memory
// (two block barriers, one more trip through the LSU / shared-memory pipe, which is the busiest unit of every c64
// kernel).  Here ONE THREAD owns a whole column: it runs the eight butterflies of the first level, parks the 64
// results in its own lane of the SM's 256 KiB tensor memory (tcgen05.st, 32x32b shape: 1 KiB = 256 columns per
// thread) and fetches them back eight at a time for the butterflies of the second level (tcgen05.ld).  No shared
// memory, no barrier between warps, and HBM sees 512 contiguous bytes per warp request (lanes = consecutive
// columns).  tools/tmem_probe.cu measured the park + fetch round trip on B200: 128 KiB per CTA in 2233 cycles
// (59 B/clk each way with one CTA per SM, 77 B/clk per SM with two), on a datapath the LSU does not share.
//
// A CTA is 128 threads = one thread per TMEM lane, 256 columns; two CTAs fill an SM's tensor memory.
// Same butterflies (c64_math.cuh) and twiddle values as c64_column.cu => bit-identical results.
#include "c64_dev.cuh"
#include "plan.h"

namespace cfft {
using namespace dev;
namespace {

__device__ __forceinline__ void tmem_st_c64(uint32_t taddr, c64 v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__double2loint(v.x)),
                 "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)), "r"(__double2hiint(v.y))
                 : "memory");
}
__device__ __forceinline__ c64 tmem_ld_c64(uint32_t taddr)
{
    int a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
    // the wait is part of the load here: the registers may be read as soon as this function returns
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(a), "+r"(b), "+r"(c), "+r"(d)::"memory");
    return mk(__hiloint2double(b, a), __hiloint2double(d, c));
}
// eight c64 (32 consecutive TMEM columns of this thread's lane) in one instruction
__device__ __forceinline__ void tmem_ld_c64x8(uint32_t taddr, c64 (&x)[8])
{
    int r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    // every destination register is an in/out operand of the wait, so no use of them can be scheduled above it
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = mk(__hiloint2double(r[4 * k + 1], r[4 * k]), __hiloint2double(r[4 * k + 3], r[4 * k + 2]));
}
__device__ __forceinline__ void tmem_st_c64x8(uint32_t taddr, const c64 (&x)[8])
{
    int r[32];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        r[4 * k] = __double2loint(x[k].x);
        r[4 * k + 1] = __double2hiint(x[k].x);
        r[4 * k + 2] = __double2loint(x[k].y);
        r[4 * k + 3] = __double2hiint(x[k].y);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t kTmemCols = 256; // per CTA: 128 lanes x 256 columns x 4 B = 128 KiB = 64 c64 per thread

struct TmemColParams {
    uint32_t n;               // transform size
    uint32_t span0;           // span of the first level of the pair
    uint32_t stride;          // span0 / 64: distance between the rows of a column
    uint32_t tiles_per_chunk; // stride / 128
    uint32_t tiles_per_row;   // n / (64 * 128)
    const c64 *tw0;           // planar twiddles of the outer level: w_k[p] at tw0[(k-1) * 8 stride + p]
    const c64 *tw1;           // planar twiddles of the inner level: w_k[p] at tw1[(k-1) * stride + p]
};

// Element (row, col) of the tile at g[row * stride + col].  Rows of the pair of levels (col_level<8, 64, ...> then
// col_level<8, 8, ...> in c64_column.cu): outer butterfly `prow` takes rows prow + 8 k and leaves output k in row
// prow + 8 brev(k); inner butterfly `b` takes rows 8 b + k and leaves output k in row 8 b + brev(k).  The inverse
// mirrors it (inputs from the bit-reversed slots, twiddles on the inputs, inner level first).
template <bool FWD>
__global__ void __launch_bounds__(128, 2) c64_tmem_column88_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, TmemColParams prm)
{
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t slot = static_cast<uint32_t>(__cvta_generic_to_shared(&tmem_base_slot));
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w owns TMEM lanes 32 w .. 32 w + 31 (a warp can only reach its own lane quadrant); thread = lane
    const uint32_t taddr = tmem_base_slot + ((uint32_t(warp) * 32u) << 16);

    const uint32_t tile = blockIdx.x;
    const uint32_t row = tile / prm.tiles_per_row;
    const uint32_t tt = tile - row * prm.tiles_per_row;
    const uint32_t chunk = tt / prm.tiles_per_chunk;
    const uint32_t col = (tt - chunk * prm.tiles_per_chunk) * 128u + threadIdx.x; // column inside the chunk
    const size_t goff = size_t(row) * prm.n + size_t(chunk) * prm.span0 + col;
    const c64 *g = src + goff;
    c64 *go = dst + goff;
    const size_t st = prm.stride;
    const size_t m0 = 8 * st, m1 = st;

    if (FWD) {
        // outer level: 8 butterflies, results parked at TMEM columns 4 * row
#pragma unroll
        for (int prow = 0; prow < 8; prow++) {
            c64 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = ld_stream(g + (prow + 8 * k) * st);
            bf8<true>(x);
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(prm.tw0 + (k - 1) * m0 + prow * st + col), x[k]);
#pragma unroll
            for (int k = 0; k < 8; k++) tmem_st_c64(taddr + 4u * uint32_t(prow + 8 * brev_c<8>(k)), x[k]);
        }
        tmem_wait_st();
        c64 w[8];
#pragma unroll
        for (int k = 1; k < 8; k++) w[k] = ld_tw(prm.tw1 + (k - 1) * m1 + col);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            c64 x[8];
            tmem_ld_c64x8(taddr + 32u * uint32_t(b), x); // rows 8 b .. 8 b + 7
            bf8<true>(x);
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(w[k], x[k]);
#pragma unroll
            for (int k = 0; k < 8; k++) st_stream(go + (8 * b + brev_c<8>(k)) * st, x[k]);
        }
    } else {
        c64 w[8];
#pragma unroll
        for (int k = 1; k < 8; k++) w[k] = ld_tw(prm.tw1 + (k - 1) * m1 + col);
        // inner level first: inputs from the bit-reversed slots, twiddles on the inputs, outputs to rows 8 b + k
#pragma unroll
        for (int b = 0; b < 8; b++) {
            c64 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = ld_stream(g + (8 * b + brev_c<8>(k)) * st);
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(w[k], x[k]);
            bf8<false>(x);
            tmem_st_c64x8(taddr + 32u * uint32_t(b), x);
        }
        tmem_wait_st();
#pragma unroll
        for (int prow = 0; prow < 8; prow++) {
            c64 x[8];
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = tmem_ld_c64(taddr + 4u * uint32_t(prow + 8 * brev_c<8>(k)));
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(prm.tw0 + (k - 1) * m0 + prow * st + col), x[k]);
            bf8<false>(x);
#pragma unroll
            for (int k = 0; k < 8; k++) st_stream(go + (prow + 8 * k) * st, x[k]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "r"(kTmemCols) : "memory");
}

} // namespace

bool tmem_column88_supported(uint32_t n, uint32_t span0)
{
    const uint32_t stride = span0 / 64;
    return span0 >= 64 * 128 && stride % 128 == 0 && n % span0 == 0;
}

// two radix-8 levels (spans span0 and span0 / 8) of `batch` transforms in one HBM pass; src == dst: in place
cudaError_t launch_c64_tmem_column88(bool inverse, const double2 *src, double2 *dst, uint64_t batch, uint32_t n, uint32_t span0,
                                     const double2 *tw0, const double2 *tw1, cudaStream_t stream)
{
    if (!tmem_column88_supported(n, span0)) return cudaErrorInvalidValue;
    TmemColParams prm;
    prm.n = n;
    prm.span0 = span0;
    prm.stride = span0 / 64;
    prm.tiles_per_chunk = prm.stride / 128;
    prm.tiles_per_row = n / (64 * 128);
    prm.tw0 = tw0;
    prm.tw1 = tw1;
    const uint64_t tiles = batch * prm.tiles_per_row;
    if (tiles > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    if (inverse) c64_tmem_column88_kernel<false><<<unsigned(tiles), 128, 0, stream>>>(src, dst, prm);
    else c64_tmem_column88_kernel<true><<<unsigned(tiles), 128, 0, stream>>>(src, dst, prm);
    count_launch();
    return cudaGetLastError();
}

} // namespace cfft

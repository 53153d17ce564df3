// ubench_tma.cu -- A/B behind the column-pass design: what does one pass of a column tile through an SM cost
// when the tile (RG = 256 rows x 8 c64 columns, row stride 256 c64 = the first pass of an n = 2^16 transform) travels
//   A. through the LSU:  LDG.128 -> [STS.128, barrier, LDS.128] x E -> STG.128        (c64_column_kernel's data path)
//   B. through TMA:      cp.async.bulk.tensor.2d (SWIZZLE_128B) -> mbarrier -> [LDS.128, STS.128 in place, barrier] x (E + 1)
//                        -> fence.proxy.async -> cp.async.bulk.tensor.2d store      (persistent CTAs, S-stage ring)
// with no arithmetic at all: the ceiling of each data path, L2-resident (32 MiB) and HBM-resident (1 GiB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tma ubench_tma.cu && ./ubench_tma
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                        \
        }                                                                                   \
    } while (0)

constexpr int RG = 256, CW = 8, ROWSTRIDE = 256; // tile geometry in c64
constexpr int TILE_BYTES = RG * CW * 16;         // 32 KiB

__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(double2 *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---- A: LSU path, one CTA per tile, 128 threads, 4 CTAs per SM -------------------------------------------------
// mapping m (the "level"): thread t, slot k -> (row, col); three different maps so that every exchange really moves data
template <int M> __device__ __forceinline__ void map_rc(int t, int k, int &row, int &col)
{
    col = t & 7;
    const int q = t >> 3; // 0 .. 15
    if (M == 0) row = q + 16 * k;                        // radix-16-like: stride 16 rows
    else if (M == 1) row = (q & 3) + 4 * k + 64 * (q >> 2); // stride 4 inside blocks of 64
    else row = 16 * q + k;                                // contiguous
}

template <int E> __global__ void __launch_bounds__(128, 4) lsu_tiles(double2 *data, int tiles_per_row, int tiles)
{
    __shared__ double2 s[RG * CW];
    const int t = threadIdx.x;
    const size_t tile = blockIdx.x % unsigned(tiles); // the grid sweeps the buffer several times (L2-resident case)
    const size_t j = tile / tiles_per_row, c0 = (tile % tiles_per_row) * CW;
    double2 *g = data + j * (size_t(RG) * ROWSTRIDE) + c0;
    double2 v[16];
    int row, col;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        map_rc<0>(t, k, row, col);
        v[k] = ldg_stream(g + size_t(row) * ROWSTRIDE + col);
    }
    if (E >= 1) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            map_rc<0>(t, k, row, col);
            s[row * CW + (col ^ (row & 7))] = v[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            map_rc<1>(t, k, row, col);
            v[k] = s[row * CW + (col ^ (row & 7))];
        }
    }
    if (E >= 2) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            map_rc<1>(t, k, row, col);
            s[row * CW + (col ^ (row & 7))] = v[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) {
            map_rc<2>(t, k, row, col);
            v[k] = s[row * CW + (col ^ (row & 7))];
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        map_rc<(E == 0 ? 0 : E)>(t, k, row, col);
        stg_stream(g + size_t(row) * ROWSTRIDE + col, v[k]);
    }
}

// ---- B: TMA path ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(smem_u32(b)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src))
                 : "memory");
}

// element (row, col) of a SWIZZLE_128B tile whose rows are 128 bytes: 16-byte chunk index XOR (row & 7)
__device__ __forceinline__ int swz(int row, int col) { return row * CW + (col ^ (row & 7)); }

// NLEV in-place levels (LDS 16, STS 16 to the same places), level-to-level barrier between the 128 consumer threads;
// WARPCOL: every level's 16 x 16 elements of a half-warp are one COLUMN (rows of one column), so warps never exchange
// with each other and the barrier is __syncwarp
template <int NLEV, int S, bool WARPCOL>
__global__ void __launch_bounds__(160, 1) tma_tiles(const __grid_constant__ CUtensorMap map, int tiles_total, int tiles_per_row, int tiles)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ uint64_t full[S], done[S];
    double2 *bufs = reinterpret_cast<double2 *>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int t = threadIdx.x;
    if (t == 0) {
        for (int i = 0; i < S; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&done[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int first = blockIdx.x, step = gridDim.x;
    const int mine = first < tiles_total ? (tiles_total - first + step - 1) / step : 0;
    if (t >= 128) { // producer warp, one lane
        if (t == 128) {
            for (int it = 0; it < mine + S; it++) {
                const int s = it % S;
                if (it >= S) { // tile it - S has been processed in place: write it back, then the buffer is free again
                    mbar_wait(&done[s], ((it / S) - 1) & 1);
                    const int tile = (first + (it - S) * step) % tiles;
                    const int j = tile / tiles_per_row, c0 = (tile % tiles_per_row) * CW;
                    tma_store_2d(&map, c0 * 2, j * RG, bufs + size_t(s) * RG * CW);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                if (it < mine) {
                    const int tile = (first + it * step) % tiles;
                    const int j = tile / tiles_per_row, c0 = (tile % tiles_per_row) * CW;
                    mbar_expect_tx(&full[s], TILE_BYTES);
                    tma_load_2d(bufs + size_t(s) * RG * CW, &map, c0 * 2, j * RG, &full[s]);
                }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
        return;
    }
    double2 v[16];
    for (int it = 0; it < mine; it++) {
        const int s = it % S;
        double2 *b = bufs + size_t(s) * RG * CW;
        mbar_wait(&full[s], (it / S) & 1);
#pragma unroll
        for (int lev = 0; lev < NLEV; lev++) {
            int row[16], col[16];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (WARPCOL) { // half-warp h owns column h of the tile: 16 threads x 16 rows, three different row maps
                    const int h = t >> 4, l = t & 15;
                    col[k] = h;
                    row[k] = lev == 0 ? l + 16 * k : (lev == 1 ? (l & 7) + 8 * ((l >> 3) + 2 * k) : (l & 7) + 8 * (16 * (l >> 3) + k));
                } else {
                    if (lev == 0) map_rc<0>(t, k, row[k], col[k]);
                    else if (lev == 1) map_rc<1>(t, k, row[k], col[k]);
                    else map_rc<2>(t, k, row[k], col[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = b[swz(row[k], col[k])];
#pragma unroll
            for (int k = 0; k < 16; k++) {
                v[k].x += 1.0; // keep the round trip alive
                b[swz(row[k], col[k])] = v[k];
            }
            if (lev + 1 < NLEV) {
                if (WARPCOL) __syncwarp();
                else asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&done[s]);
    }
}

// ---- C: contiguous tiles (the shape of the N = 2048 kernel, the 256-point rows pass and the generic register kernel): a CTA of
// 128 threads owns 2048 contiguous c64; LDG.128 in, E exchanges through shared memory, then out either by STG.128 (which
// ncu shows costing TWO data-pipe wavefronts per 128 bytes, l1tex__data_pipe_lsu_wavefronts_mem_lgds) or by STS.128 into the
// tile + one cp.async.bulk (TMA, 1-D) per warp of 8 KiB, which leaves the LSU data pipe out of the store.
template <int E, bool BULK> __global__ void __launch_bounds__(128, 4) contig_tiles(double2 *data, int tiles)
{
    __shared__ __align__(128) double2 s[2048];
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    double2 *g = data + size_t(blockIdx.x % unsigned(tiles)) * 2048;
    double2 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ldg_stream(g + t + 128 * k);
#pragma unroll
    for (int e = 0; e < E; e++) {
        double2 *blk = s + 256 * (t >> 4); // the 16 x 16 transpose of base256 (c64_dev.cuh): XOR swizzle, conflict-free
        const int l16 = t & 15;
#pragma unroll
        for (int k = 0; k < 16; k++) blk[16 * l16 + (k ^ l16)] = v[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = blk[16 * k + (l16 ^ k)];
        __syncwarp();
    }
    if (!BULK) {
#pragma unroll
        for (int k = 0; k < 16; k++) stg_stream(g + t + 128 * k, v[k]);
    } else {
        // warp w stores its own 512 contiguous elements (8 KiB): natural order, lanes on consecutive c64
#pragma unroll
        for (int k = 0; k < 16; k++) s[512 * w + l + 32 * k] = v[k];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (l == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 8192;" ::"l"(g + 512 * w), "r"(smem_u32(s + 512 * w)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
}

template <int E, bool BULK> static void run_contig(const char *name, double2 *d, size_t bytes, int reps, int sweeps)
{
    const int tiles = int(bytes / TILE_BYTES);
    const float ms = time_ms(reps, [&] { contig_tiles<E, BULK><<<tiles * sweeps, 128>>>(d, tiles); });
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    printf("%-44s                   %8.3f ms  %7.0f GB/s (read + write)\n", name, ms, 2.0 * bytes * sweeps / ms / 1e6);
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static float time_ms(int reps, const std::function<void()> &fn)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    fn();
    fn();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) fn();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ms / reps;
}

template <int NLEV, int S, bool WARPCOL> static void run_tma(const char *name, const CUtensorMap &map, size_t bytes, int sms, int ctas_per_sm, int reps, int sweeps)
{
    const int tiles = int(bytes / TILE_BYTES), tiles_per_row = ROWSTRIDE / CW;
    const size_t smem = size_t(S) * TILE_BYTES + 1024;
    auto k = tma_tiles<NLEV, S, WARPCOL>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 160, smem));
    if (occ < ctas_per_sm) {
        printf("%-44s skipped (occupancy %d < %d)\n", name, occ, ctas_per_sm);
        return;
    }
    const int grid = sms * ctas_per_sm;
    const float ms = time_ms(reps, [&] { k<<<grid, 160, smem>>>(map, tiles * sweeps, tiles_per_row, tiles); });
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    printf("%-44s S=%d x %d CTA/SM  %8.3f ms  %7.0f GB/s (read + write)\n", name, S, ctas_per_sm, ms, 2.0 * bytes * sweeps / ms / 1e6);
}

template <int E> static void run_lsu(const char *name, double2 *d, size_t bytes, int reps, int sweeps)
{
    const int tiles = int(bytes / TILE_BYTES), tiles_per_row = ROWSTRIDE / CW;
    const float ms = time_ms(reps, [&] { lsu_tiles<E><<<tiles * sweeps, 128>>>(d, tiles_per_row, tiles); });
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    printf("%-44s                   %8.3f ms  %7.0f GB/s (read + write)\n", name, ms, 2.0 * bytes * sweeps / ms / 1e6);
}

int main()
{
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void **>(&encode), cudaEnableDefault, &qres));
    if (!encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    for (size_t mib : {32, 1024}) {
        const size_t bytes = mib << 20;
        double2 *d = nullptr;
        CK(cudaMalloc(&d, bytes));
        CK(cudaMemset(d, 0, bytes));
        // the buffer as a 2-D tensor of doubles: inner dimension = one row of 256 c64 = 512 doubles, outer = all rows
        CUtensorMap map;
        const cuuint64_t dims[2] = {2 * ROWSTRIDE, bytes / (ROWSTRIDE * 16)};
        const cuuint64_t strides[1] = {ROWSTRIDE * 16};
        const cuuint32_t box[2] = {2 * CW, RG};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            printf("cuTensorMapEncodeTiled failed: %d\n", int(r));
            return 1;
        }
        const int reps = 10, sweeps = mib <= 64 ? 64 : 1; // L2-resident: 64 sweeps per launch so that launch overhead vanishes
        printf("---- buffer %zu MiB (%s), tiles of %d rows x %d c64, row stride %d c64 ----\n", mib, mib <= 64 ? "L2-resident" : "HBM", RG, CW, ROWSTRIDE);
        run_lsu<0>("LSU: LDG -> STG", d, bytes, reps, sweeps);
        run_lsu<1>("LSU: LDG, 1 exchange, STG", d, bytes, reps, sweeps);
        run_lsu<2>("LSU: LDG, 2 exchanges, STG", d, bytes, reps, sweeps);
        run_tma<1, 3, false>("TMA: 1 in-place level", map, bytes, sms, 1, reps, sweeps);
        run_tma<1, 3, false>("TMA: 1 in-place level", map, bytes, sms, 2, reps, sweeps);
        run_tma<1, 2, false>("TMA: 1 in-place level", map, bytes, sms, 3, reps, sweeps);
        run_tma<2, 3, false>("TMA: 2 in-place levels, block barrier", map, bytes, sms, 2, reps, sweeps);
        run_tma<3, 3, false>("TMA: 3 in-place levels, block barrier", map, bytes, sms, 2, reps, sweeps);
        run_tma<3, 2, false>("TMA: 3 in-place levels, block barrier", map, bytes, sms, 3, reps, sweeps);
        run_tma<3, 3, true>("TMA: 3 in-place levels, half-warp columns", map, bytes, sms, 2, reps, sweeps);
        run_tma<3, 2, true>("TMA: 3 in-place levels, half-warp columns", map, bytes, sms, 3, reps, sweeps);
        run_tma<3, 6, true>("TMA: 3 in-place levels, half-warp columns", map, bytes, sms, 1, reps, sweeps);
        run_contig<0, false>("contig: LDG -> STG", d, bytes, reps, sweeps);
        run_contig<0, true>("contig: LDG -> STS + bulk store", d, bytes, reps, sweeps);
        run_contig<1, false>("contig: LDG, 1 exchange, STG", d, bytes, reps, sweeps);
        run_contig<1, true>("contig: LDG, 1 exchange, STS + bulk store", d, bytes, reps, sweeps);
        run_contig<2, false>("contig: LDG, 2 exchanges, STG", d, bytes, reps, sweeps);
        run_contig<2, true>("contig: LDG, 2 exchanges, STS + bulk store", d, bytes, reps, sweeps);
        run_contig<3, false>("contig: LDG, 3 exchanges, STG", d, bytes, reps, sweeps);
        run_contig<3, true>("contig: LDG, 3 exchanges, STS + bulk store", d, bytes, reps, sweeps);
        CK(cudaFree(d));
    }
    return 0;
}

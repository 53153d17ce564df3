// c64_ord16.cu -- register-resident kernels for standard-order Dif16 plans below and above the
// 256-point case that c64_fast.cu covers:
//   ordered::Plan::new(n, UserProvided(Dif16)) / Measure,  n = 16, 32, 64, 128, 512, 1024
//   unordered::Plan with base_n == n and base_algo == Dif16 (the same transform, src/unordered.rs:561-564)
//
// Stage schedule of the reference for these sizes (src/dif16.rs:449-623 stockham_core_generic,
// :649-827 stockham_dif16_end, recursion as src/dif4.rs:285-303):
//   n = 16                   :  terminal radix-16 only
//   n = 16 R3  (R3 = 2, 4, 8):  radix-16 at stride 1 with twiddles, terminal radix-R3 at stride 16
//   n = 256 R3 (R3 = 2, 4)   :  radix-16 at stride 1, radix-16 at stride 16 (both with twiddles),
//                               terminal radix-R3 at stride 256
// with  y[q + s(16 p + k)] = w_n^{k p s} * DFT16(x[q + s(p + m k)])_k,  m = n / (16 s),
// and the inverse = same schedule, conjugated table, mirrored butterflies (src/fft_simd.rs:113-120).
// Same butterflies (c64_math.cuh), same table values (the planar half of init_wt, src/fft_simd.rs:311-316)
// => bit-identical to the reference / the exact tile kernel; checked against the oracle in the tests.
//
// Mapping: n/16 threads per transform, 16 c64 per thread, every stage in registers, exchanges through
// XOR-swizzled shared memory.  n >= 512 reads and writes HBM directly (lanes on consecutive c64);
// n <= 128 would touch only 32..128 contiguous bytes per request that way, so a CTA first stages 2048
// contiguous elements (64 / 32 / 16 transforms) through shared memory with fully coalesced accesses.
#include "c64_dev.cuh"
#include "plan.h"

namespace cfft {
using namespace dev;
namespace {

constexpr int kNT = 128; // threads per CTA

// ---- n = 512, 1024 ---------------------------------------------------------------------------
template <int N, bool FWD>
__global__ void __launch_bounds__(kNT, 4)
c64_ord16_mid_kernel(c64 *__restrict__ data, uint64_t batch, const c64 *__restrict__ tw)
{
    constexpr int TPR = N / 16, ROWS = kNT / TPR, R3 = N / 256;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int row = threadIdx.x / TPR, t = threadIdx.x % TPR;
    const uint64_t grow = uint64_t(blockIdx.x) * ROWS + row;
    const bool active = grow < batch; // warp-uniform (TPR >= 32); inactive rows recompute row 0 and store nothing
    c64 *g = data + (active ? grow : 0) * N;
    c64 *s = reinterpret_cast<c64 *>(smem_raw) + row * N;
    c64 v[16];
    auto row_sync = [] { // the TPR threads of a row
        if (TPR == 32) __syncwarp();
        else __syncthreads();
    };

    // stride 1: x[p + TPR k] -> y[16 p + k], p = t
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = ld_stream(g + t + TPR * k);
    bf16<FWD>(v);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw + t + TPR * k), v[k]);
#pragma unroll
    for (int k = 0; k < 16; k++) s[16 * t + (k ^ (t & 15))] = v[k];
    row_sync();

    // stride 16: y[q + 16 (p2 + R3 k)] -> z[q + 16 (16 p2 + k)], twiddle w_n^{16 k p2}
    const int q = t & 15, p2 = t >> 4;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int P = p2 + R3 * k;
        v[k] = s[16 * P + (q ^ (P & 15))];
    }
    bf16<FWD>(v);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw + 16 * p2 + TPR * k), v[k]);
    row_sync(); // the whole row has been read before it is overwritten in natural order
#pragma unroll
    for (int k = 0; k < 16; k++) s[256 * p2 + 16 * k + q] = v[k];
    row_sync();

    // terminal radix R3 at stride 256, 16 / R3 butterflies per thread, in place -> HBM
#pragma unroll
    for (int i = 0; i < 16 / R3; i++)
#pragma unroll
        for (int k = 0; k < R3; k++) v[i * R3 + k] = s[t + TPR * i + 256 * k];
#pragma unroll
    for (int i = 0; i < 16 / R3; i++) bfR<R3, FWD>(&v[i * R3]);
    if (active) {
#pragma unroll
        for (int i = 0; i < 16 / R3; i++)
#pragma unroll
            for (int k = 0; k < R3; k++) st_stream(g + t + TPR * i + 256 * k, v[i * R3 + k]);
    }
}

// ---- n = 32, 64, 128 -------------------------------------------------------------------------
template <int N, bool FWD>
__global__ void __launch_bounds__(kNT, 4)
c64_ord16_small_kernel(c64 *__restrict__ data, uint64_t total, const c64 *__restrict__ tw)
{
    constexpr int R3 = N / 16;   // threads per transform = radix of the terminal pass
    constexpr int TILE = 2048;   // c64 per CTA = 16 per thread
    constexpr int G = 8 / R3;    // rows that share one 8-lane shared-memory phase
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *s = reinterpret_cast<c64 *>(smem_raw);
    const uint64_t base = uint64_t(blockIdx.x) * TILE;
    c64 *g = data + base;
    const uint32_t valid = total - base < TILE ? uint32_t(total - base) : TILE;
    // natural-order layout: position pos of row r lives at r N + (pos ^ ((r mod G) R3)), which keeps
    // lanes (r, p) that read p + R3 k in different 16-byte bank groups
    auto swz = [](int e) { return e ^ (((e / N) & (G - 1)) * R3); };
    c64 v[16];

#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t idx = threadIdx.x + kNT * j;
        v[j] = idx < valid ? ld_stream(g + idx) : mk(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < 16; j++) s[swz(int(threadIdx.x) + kNT * j)] = v[j];
    __syncthreads();

    const int L = threadIdx.x, row = L / R3, p = L % R3;
    const int rs = (row & (G - 1)) * R3;
    c64 *sr = s + row * N;
    // stride 1: x[p + R3 k] -> y[16 p + k]; rows are owned by R3 <= 8 lanes of one warp
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = sr[(p + R3 * k) ^ rs];
    bf16<FWD>(v);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw + p + R3 * k), v[k]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) s[16 * L + (k ^ (L & 15))] = v[k]; // 16 L = row N + 16 p
    __syncwarp();
    // terminal radix R3 at stride 16: thread p owns q = p + R3 i
#pragma unroll
    for (int i = 0; i < 16 / R3; i++)
#pragma unroll
        for (int k = 0; k < R3; k++) {
            const int Lk = row * R3 + k;
            v[i * R3 + k] = s[16 * Lk + ((p + R3 * i) ^ (Lk & 15))];
        }
#pragma unroll
    for (int i = 0; i < 16 / R3; i++) bfR<R3, FWD>(&v[i * R3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16 / R3; i++)
#pragma unroll
        for (int k = 0; k < R3; k++) sr[(p + R3 * i + 16 * k) ^ rs] = v[i * R3 + k];
    __syncthreads();

#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = s[swz(int(threadIdx.x) + kNT * j)];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t idx = threadIdx.x + kNT * j;
        if (idx < valid) st_stream(g + idx, v[j]);
    }
}

// ---- n = 16: the whole transform is the terminal radix-16 butterfly (src/dif16.rs:649-827) ------
template <bool FWD>
__global__ void __launch_bounds__(kNT, 4)
c64_ord16_n16_kernel(c64 *__restrict__ data, uint64_t total)
{
    constexpr int TILE = 2048; // 128 transforms per CTA, one per thread
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *s = reinterpret_cast<c64 *>(smem_raw);
    const uint64_t base = uint64_t(blockIdx.x) * TILE;
    c64 *g = data + base;
    const uint32_t valid = total - base < TILE ? uint32_t(total - base) : TILE;
    auto swz = [](int e) { return (e & ~15) | ((e ^ (e >> 4)) & 15); }; // element k of transform L at 16 L + (k ^ (L & 15))
    c64 v[16];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t idx = threadIdx.x + kNT * j;
        v[j] = idx < valid ? ld_stream(g + idx) : mk(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < 16; j++) s[swz(int(threadIdx.x) + kNT * j)] = v[j];
    __syncthreads();
    const int L = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = s[16 * L + (k ^ (L & 15))];
    bf16<FWD>(v);
#pragma unroll
    for (int k = 0; k < 16; k++) s[16 * L + (k ^ (L & 15))] = v[k];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; j++) v[j] = s[swz(int(threadIdx.x) + kNT * j)];
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const uint32_t idx = threadIdx.x + kNT * j;
        if (idx < valid) st_stream(g + idx, v[j]);
    }
}

cudaError_t launch_n16(bool inverse, c64 *data, uint64_t batch, cudaStream_t st)
{
    const uint64_t total = batch * 16;
    const size_t smem = 2048 * sizeof(c64);
    const unsigned ctas = unsigned((total + 2047) / 2048);
    if (inverse) c64_ord16_n16_kernel<false><<<ctas, kNT, smem, st>>>(data, total);
    else c64_ord16_n16_kernel<true><<<ctas, kNT, smem, st>>>(data, total);
    count_launch();
    return cudaGetLastError();
}

template <int N> cudaError_t launch_mid(bool inverse, c64 *data, uint64_t batch, const c64 *tw, cudaStream_t st)
{
    constexpr int ROWS = kNT / (N / 16);
    const size_t smem = size_t(ROWS) * N * sizeof(c64); // 32 KiB
    const unsigned ctas = unsigned((batch + ROWS - 1) / ROWS);
    if (inverse) c64_ord16_mid_kernel<N, false><<<ctas, kNT, smem, st>>>(data, batch, tw);
    else c64_ord16_mid_kernel<N, true><<<ctas, kNT, smem, st>>>(data, batch, tw);
    count_launch();
    return cudaGetLastError();
}

template <int N> cudaError_t launch_small(bool inverse, c64 *data, uint64_t batch, const c64 *tw, cudaStream_t st)
{
    const uint64_t total = batch * N;
    const size_t smem = 2048 * sizeof(c64);
    const unsigned ctas = unsigned((total + 2047) / 2048);
    if (inverse) c64_ord16_small_kernel<N, false><<<ctas, kNT, smem, st>>>(data, total, tw);
    else c64_ord16_small_kernel<N, true><<<ctas, kNT, smem, st>>>(data, total, tw);
    count_launch();
    return cudaGetLastError();
}

} // namespace

bool ord16_supported(uint64_t n, int algo)
{
    return algo == 6 /* Dif16 */ && (n == 16 || n == 32 || n == 64 || n == 128 || n == 512 || n == 1024);
}

// tw: the plan's table for the direction; its first n entries are the planar half of init_wt(16, n)
cudaError_t launch_c64_ord16(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t st)
{
    if (batch == 0) return cudaSuccess;
    const c64 *tw = plan->d_tw[inverse ? 1 : 0];
    switch (plan->n) {
    case 16: return launch_n16(inverse, data, batch, st);
    case 32: return launch_small<32>(inverse, data, batch, tw, st);
    case 64: return launch_small<64>(inverse, data, batch, tw, st);
    case 128: return launch_small<128>(inverse, data, batch, tw, st);
    case 512: return launch_mid<512>(inverse, data, batch, tw, st);
    case 1024: return launch_mid<1024>(inverse, data, batch, tw, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cfft

// c64_fast.cu -- register-resident c64 kernels for unordered plans with base (Dif16, 256):
//   n = 256 * R1 * R2,  R1, R2 in {1, 2, 4, 8}  ->  n = 256 .. 8192   (TFHE polynomial sizes)
//   n > 8192: the levels run as column passes (c64_column.cu), then this file's N = 256 kernel
//   runs the base FFTs on contiguous rows.
//
// Same butterflies, same twiddle values and same permuted output order as the reference for
// Method::UserProvided { base_algo: Dif16, base_n: 256 }, so the result is bit-identical to
// concrete-fft's (checked against the oracle in tests/test_gpu_c64.py):
//   unordered levels  fwd_process_x{2,4,8} / inv_process_x{2,4,8}   src/unordered.rs:222-293
//   base FFT          Dif16 stockham_core (s = 1) + stockham_dif16_end   src/dif16.rs:449-827
//
// Mapping: n/16 threads per transform, 16 c64 (64 registers) per thread.  Every stage is one
// radix-16 butterfly (or 16/r radix-r butterflies) per thread, entirely in registers.
//   * HBM -> registers with 128-bit loads, lanes on consecutive c64 (512 B per warp request);
//     registers -> HBM the same way: a transform moves exactly 2 * 16 * n bytes.
//   * Exchanges between levels go through shared memory in natural order (conflict-free).
//   * The 16x16 transpose between the two radix-16 passes of a 256-point base FFT stays inside
//     one half-warp: XOR-swizzled shared memory + __syncwarp, no block barrier.
//   * Twiddles come from plan tables re-laid out planar (w_k[p], lanes on consecutive p) so
//     each request is one or two 128 B lines; they stay L1-resident.
#include <cooperative_groups.h>

#include <cstdlib>
#include <mutex>

#include "c64_fast_kernels.cuh"

namespace cfft {
using namespace dev;
using namespace fastk;
constexpr uint64_t kWorkspaceChunkBytes = uint64_t{256} << 20;
namespace {

// ---- n = 8192 / 16384: one transform per thread-block CLUSTER --------------------------------
// A 128 / 256 KiB transform does not fit one SM with room for a second CTA, so the fused kernel above
// runs one CTA per SM and its load / compute / store phases cannot overlap.  Here CSZ = 2 / 4 CTAs of
// 256 threads and 64 KiB each own one transform: the first level (radix 8) reads HBM and scatters
// its eight output chunks into the owning CTA's shared memory over DSMEM (about 30 GB/s per SM on
// B200, tools/dsmem_probe.cu; half / three quarters of the data crosses once), then every CTA
// finishes its 4096 contiguous elements locally (second level + 256-point base FFTs) exactly like
// the single-CTA kernel.  Two CTAs of different clusters share an SM, so HBM, DSMEM and FP64 phases
// overlap.  Same butterflies and twiddles => same bits.
namespace cg = cooperative_groups;

template <int N, int CSZ, int R2, bool FWD>
__global__ void __launch_bounds__(256, 2)
c64_cluster_kernel(c64 *__restrict__ data, FastTables tb)
{
    constexpr int NL = N / CSZ;   // elements finished by one CTA (4096)
    constexpr int M1 = N / 8;     // chunk size after the radix-8 level
    constexpr int CPC = 8 / CSZ;  // chunks owned per CTA
    constexpr int PPC = M1 / CSZ; // first-level butterflies (values of p) per CTA = 512
    static_assert(NL == 4096 && PPC == 512 && R2 * 256 == M1, "cluster layout");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *s = reinterpret_cast<c64 *>(smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = int(cluster.block_rank());
    const int t = threadIdx.x;
    c64 *g = data + size_t(blockIdx.x / CSZ) * N;
    c64 v[16];
    const int blk = t / 16, lane16 = t % 16;
    // (An L2 prefetch of the successor cluster's transform, the trick that gives the single-CTA kernel +23 % at this size,
    // changes nothing here -- profiles/r2d_prefetch_bulkstore_twsplit_ab.txt: two CTAs per SM already overlap their loads.)

    if (FWD) {
        cluster.sync(); // every CTA of the cluster is resident before anyone writes remote shared memory
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 8; k++) v[j * 8 + k] = ld_stream(g + rank * PPC + t + 256 * j + M1 * k);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            c64 *x = &v[j * 8];
            const int p = rank * PPC + t + 256 * j;
            bf8<true>(x);
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(tb.top1 + (k - 1) * M1 + p), x[k]);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int c = brev_c<8>(k);
                c64 *dst = cluster.map_shared_rank(s, unsigned(c / CPC));
                dst[(c % CPC) * M1 + p] = x[k];
            }
        }
        cluster.sync();
        level<R2, M1, 256, true, false, false>(s, s, tb.top2, t, v);
        __syncthreads();
        base256<true, false, true>(s + blk * 256, s + blk * 256, g + rank * NL + blk * 256, tb.base, lane16, v);
    } else {
        base256<false, true, false>(g + rank * NL + blk * 256, s + blk * 256, s + blk * 256, tb.base, lane16, v);
        __syncthreads();
        level<R2, M1, 256, false, false, false>(s, s, tb.top2, t, v);
        cluster.sync();
#pragma unroll
        for (int j = 0; j < 2; j++) {
            c64 *x = &v[j * 8];
            const int p = rank * PPC + t + 256 * j;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int c = brev_c<8>(k);
                const c64 *src = cluster.map_shared_rank(s, unsigned(c / CPC));
                x[k] = src[(c % CPC) * M1 + p];
            }
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = cmul(ld_tw(tb.top1 + (k - 1) * M1 + p), x[k]);
            bf8<false>(x);
#pragma unroll
            for (int k = 0; k < 8; k++) st_stream(g + p + M1 * k, x[k]);
        }
        cluster.sync(); // peers may still be reading this CTA's shared memory
    }
}

template <int N, int CSZ, int R2>
cudaError_t launch_cluster(bool inverse, c64 *data, uint64_t batch, const FastTables &tb, cudaStream_t stream)
{
    constexpr size_t smem = size_t(N / CSZ) * sizeof(c64);
    auto fk = c64_cluster_kernel<N, CSZ, R2, true>;
    auto ik = c64_cluster_kernel<N, CSZ, R2, false>;
    static thread_local int configured_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device != dev) {
        cudaError_t e = cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ik, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) return e;
        configured_device = dev;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(unsigned(batch * CSZ));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = CSZ;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaError_t e = inverse ? cudaLaunchKernelEx(&cfg, ik, data, tb) : cudaLaunchKernelEx(&cfg, fk, data, tb);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int N, int R1, int R2, bool STD = false>
cudaError_t launch_cfg(bool inverse, c64 *data, uint64_t batch, const FastTables &tb, cudaStream_t stream, uint64_t row_stride = N)
{
    using Cfg = FastCfg<N>;
    const size_t smem = size_t(Cfg::ROWS) * N * sizeof(c64);
    const uint64_t ctas = (batch + Cfg::ROWS - 1) / Cfg::ROWS;
    auto fwd_k = c64_fast_b256_kernel<N, R1, R2, true, STD>;
    auto inv_k = c64_fast_b256_kernel<N, R1, R2, false, STD>;
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_device != dev) {
            cudaError_t e = cudaFuncSetAttribute(fwd_k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(inv_k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            configured_device = dev;
        }
    }
    BatchIo<false, false> bio = plain_batch(data, data, row_stride, row_stride);
    // L2 prefetch of the successor CTA's rows (c64_fast_kernels.cuh): CFFT_B200_FAST_PREFETCH = waves ahead (0 = off)
    static const int env_pf = [] { const char *e = getenv("CFFT_B200_FAST_PREFETCH"); return e ? atoi(e) : -1; }();
    // measured (profiles/r2c_fast_prefetch.txt): n = 8192 3.74 / 3.78 -> 4.61 / 4.57 TB/s, n = 4096 fwd 5.31 -> 6.15 (inv 6.11 -> 5.99:
    // left off), n = 2048 6.66 / 6.84 -> 6.90 / 6.88
    // standard-order variant (profiles/r2l_std_one_exchange_ab.txt): on in both directions at every size (n = 4096 fwd 5.24 -> 5.87)
    // n = 1024: fwd 6.51 -> 6.84 TB/s (profiles/r2o_small_n_prefetch.txt); n <= 512: no effect, left off
    const bool on = N >= 1024 && (STD || N != 4096 || !inverse);
    const int waves = env_pf >= 0 ? env_pf : (on ? 1 : 0);
    if (waves > 0 && N >= 512) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        bio.ahead = uint32_t(sms) * uint32_t(Cfg::MINB * Cfg::ROWS * waves);
    }
    if (inverse) inv_k<<<unsigned(ctas), Cfg::NT, smem, stream>>>(bio, batch, tb);
    else fwd_k<<<unsigned(ctas), Cfg::NT, smem, stream>>>(bio, batch, tb);
    count_launch();
    return cudaGetLastError();
}

// ALLOW_MULTI = false (n = 8192: 512 threads fill the register file at 128 registers each): one term per launch
template <int N, int R1, int R2, bool ALLOW_MULTI = true>
cudaError_t launch_fused_mul(const c64 *a, const c64 *b, c64 *out, uint64_t batch, uint32_t kterms, uint64_t b_row_stride,
                             const FastTables &tf, const FastTables &ti, cudaStream_t stream, uint32_t chain = 0)
{
    using Cfg = FastCfg<N>;
    const size_t smem = size_t(Cfg::ROWS) * N * sizeof(c64);
    const uint64_t ctas = (batch + Cfg::ROWS - 1) / Cfg::ROWS;
    auto k1 = c64_fwd_mul_inv_kernel<N, R1, R2, false>;
    auto km = c64_fwd_mul_inv_kernel<N, R1, R2, ALLOW_MULTI>;
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_device != dev) {
            cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e == cudaSuccess) e = cudaFuncSetAttribute(km, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            configured_device = dev;
        }
    }
    static const int env_flags = [] { const char *e = getenv("CFFT_B200_FUSED_MUL_FLAGS"); return e ? atoi(e) : 3; }();
    uint32_t flags = uint32_t(env_flags) & 3;
    // bits 8..: L2 prefetch distance in rows = one wave of resident CTAs (c64_fast_kernels.cuh); CFFT_B200_FUSED_MUL_PREFETCH=0: off
    static const int env_pf = [] { const char *e = getenv("CFFT_B200_FUSED_MUL_PREFETCH"); return e ? atoi(e) : 1; }();
    auto ahead_bits = [&](int minb) -> uint32_t {
        if (env_pf <= 0 || N < 2048) return 0u;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return uint32_t(sms * minb * Cfg::ROWS * env_pf) << 8;
    };
    const uint32_t pf1 = ahead_bits(FusedMulCfg<N, false>::MINB), pfm = ahead_bits(FusedMulCfg<N, ALLOW_MULTI>::MINB);
    if (chain) {
        // cfft_c64_fwd_mul_add: ONE term per row (rows kterms * N apart), Fourier-domain result stored (bit 4) on top of
        // what `out` holds (bit 3) -- no inverse
        k1<<<unsigned(ctas), Cfg::NT, smem, stream>>>(plain_batch(a, nullptr, N, 0), b, plain_batch(nullptr, out, 0, N), batch, kterms, b_row_stride, tf, ti, (flags & 1) | chain | pf1);
        count_launch();
        return cudaGetLastError();
    }
    if (!ALLOW_MULTI && kterms > 1) {
        // one launch per term: out holds the Fourier-domain partial sum between launches, the last launch inverts it
        for (uint32_t k = 0; k < kterms; k++) {
            const uint32_t f = (flags & 1) | (k > 0 ? 8u : 0u) | (k + 1 < kterms ? 16u : 0u) | pf1;
            k1<<<unsigned(ctas), Cfg::NT, smem, stream>>>(plain_batch(a + uint64_t(k) * N, nullptr, N, 0), b + uint64_t(k) * N, plain_batch(nullptr, out, 0, N), batch, kterms, b_row_stride, tf, ti, f);
            count_launch();
        }
        return cudaGetLastError();
    }
    const BatchIo<false, false> ain = plain_batch(a, nullptr, N, 0), oout = plain_batch(nullptr, out, 0, N);
    if (kterms == 1 && !(env_flags & 4)) k1<<<unsigned(ctas), Cfg::NT, smem, stream>>>(ain, b, oout, batch, kterms, b_row_stride, tf, ti, flags | pf1);
    else km<<<unsigned(ctas), Cfg::NT, smem, stream>>>(ain, b, oout, batch, kterms, b_row_stride, tf, ti, flags | pfm);
    count_launch();
    return cudaGetLastError();
}

// out[r][i] = a[r * a_row_stride + i] * b[r * b_row_stride + i] (+ out[r][i] when ACC): the point-wise step of the
// composed path below, with the operand b optionally shared by every row (stride 0).
template <bool ACC>
__global__ void __launch_bounds__(256)
c64_pointwise_rows_kernel(c64 *__restrict__ out, const c64 *__restrict__ a, const c64 *__restrict__ b, uint32_t n, uint64_t rows,
                          uint64_t b_row_stride)
{
    const uint64_t total = rows * n;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t rr = i / n;
        const uint32_t c = uint32_t(i - rr * n);
        c64 p = cmul_nc(a[i], b[rr * b_row_stride + c]);
        if (ACC) p = cadd(out[i], p);
        out[i] = p;
    }
}

// ---- standard-order ("ordered") transforms above the reference's 2^10 cap ---------------------
// X_i of an n = 256 M transform sits, in the unordered layout, at row c = bitrev_L(i mod M),
// column i / M (src/unordered.rs:1046-1051 with base_n = 256).  These kernels run the 256-point
// base FFTs on TW rows whose `lo = i mod M` values are consecutive and move the tile between row
// order and standard order through a padded shared-memory transpose, so both the row side and the
// standard-order side of the pass are 256-byte-coalesced.
template <int TW, bool FWD>
__global__ void __launch_bounds__(16 * TW, TW == 16 ? 2 : 4)
c64_rows256_std_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, uint32_t n, uint32_t logm,
                       const c64 *__restrict__ tw_base)
{
    constexpr int PITCH = 257; // c64 per staged row: lanes that differ in `d` hit different banks
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *s = reinterpret_cast<c64 *>(smem_raw);
    const uint32_t m = 1u << logm;
    const uint32_t tiles_per_row = m / TW;
    const uint32_t b = blockIdx.x / tiles_per_row;
    const uint32_t lo0 = (blockIdx.x - b * tiles_per_row) * TW;
    const int d = threadIdx.x >> 4, lane16 = threadIdx.x & 15;
    const uint32_t c = __brev(lo0 + uint32_t(d)) >> (32 - logm); // row that holds lo = lo0 + d
    const size_t row_base = size_t(b) * n;
    c64 v[16];
    if (FWD) {
        // rows (unordered layout) -> base FFT -> staged natural order -> standard-order scatter
        base256<true, true, false>(src + row_base + size_t(c) * 256, s + d * PITCH, s + d * PITCH, tw_base, lane16, v);
        __syncthreads();
        const int dd = threadIdx.x % TW, h0 = threadIdx.x / TW;
#pragma unroll 4
        for (int hi = h0; hi < 256; hi += 16 * TW / TW)
            st_stream(dst + row_base + size_t(hi) * m + lo0 + dd, s[dd * PITCH + hi]);
    } else {
        const int dd = threadIdx.x % TW, h0 = threadIdx.x / TW;
        // all 16 gathers of a thread in flight before the first shared-memory store (these are 256-byte
        // segments m c64 apart: latency, not bandwidth, is what a shorter batch would expose)
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = ld_stream(src + row_base + size_t(h0 + 16 * i) * m + lo0 + dd);
#pragma unroll
        for (int i = 0; i < 16; i++) s[dd * PITCH + h0 + 16 * i] = v[i];
        __syncthreads();
        base256<false, false, true>(s + d * PITCH, s + d * PITCH, dst + row_base + size_t(c) * 256, tw_base, lane16, v);
    }
}

// TW = 16, round 2: the same pass with the un-permutation folded into the base FFT's OWN exchange instead of a second trip
// through shared memory.  The 256-point FFT is radix-16 (pass 1, x[p + 16k]) -> 16 x 16 transpose -> radix-16 (pass 2,
// y[j + 16k']), and nothing says the two passes of a row must run on the same half-warp.  Forward: pass 1 runs with
// half-warp = row d, lane = p (row-side loads coalesced); pass 2 runs with half-warp = j, lane = row d, so the sixteen
// lanes of a half-warp finish with X[hi = j + 16k''] of sixteen CONSECUTIVE lo and store 256-byte runs of standard order
// straight from registers.  Inverse: the mirror image (standard-order gathers with lanes on lo, pass 2 on rows).  The
// transpose between the passes now crosses rows, so it is one block barrier instead of a half-warp one; shared memory sees
// each element once in and once out (LSU passes per element 7.9 -> 5.9).  Layout: element (row d, p, k)
// at d * 257 + 16 p + (k ^ p): lanes varying p (pass 1) or d (pass 2) both fall into eight different 16-byte bank groups.
template <bool FWD>
__global__ void __launch_bounds__(256, 2)
c64_rows256_std16_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, uint32_t n, uint32_t logm, const c64 *__restrict__ tw_base)
{
    constexpr int TW = 16, PITCH = 257;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *s = reinterpret_cast<c64 *>(smem_raw);
    const uint32_t m = 1u << logm;
    const uint32_t tiles_per_row = m / TW;
    const uint32_t b = blockIdx.x / tiles_per_row;
    const uint32_t lo0 = (blockIdx.x - b * tiles_per_row) * TW;
    const int hw = threadIdx.x >> 4, lane = threadIdx.x & 15;
    const size_t row_base = size_t(b) * n;
    c64 v[16];
    if (FWD) {
        {   // pass 1: half-warp = row d, lane = p                                   src/dif16.rs:449-623
            const int d = hw, p = lane;
            const c64 *g = src + row_base + size_t(__brev(lo0 + uint32_t(d)) >> (32 - logm)) * 256; // the row that holds lo = lo0 + d
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ld_stream(g + p + 16 * k);
            bf16<true>(v);
#pragma unroll
            for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw_base + p + 16 * k), v[k]);
#pragma unroll
            for (int k = 0; k < 16; k++) s[d * PITCH + 16 * p + (k ^ p)] = v[k];
        }
        __syncthreads();
        {   // pass 2: half-warp = j, lane = row d                                   src/dif16.rs:649-827
            const int j = hw, d = lane;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = s[d * PITCH + 16 * k + (j ^ k)];
            bf16<true>(v);
            c64 *o = dst + row_base + lo0 + d;
#pragma unroll
            for (int k = 0; k < 16; k++) st_stream(o + size_t(j + 16 * k) * m, v[k]); // X[hi = j + 16k] at hi * M + lo
        }
    } else {
        {   // pass 1 on standard-order input: half-warp = p, lane = row d (gathers of 256-byte runs M c64 apart)
            const int p = hw, d = lane;
            const c64 *g = src + row_base + lo0 + d;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = ld_stream(g + size_t(p + 16 * k) * m);
            bf16<false>(v);
#pragma unroll
            for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw_base + p + 16 * k), v[k]);
#pragma unroll
            for (int k = 0; k < 16; k++) s[d * PITCH + 16 * p + (k ^ p)] = v[k];
        }
        __syncthreads();
        {   // pass 2: half-warp = row d, lane = j, rows written back in the unordered layout
            const int d = hw, j = lane;
#pragma unroll
            for (int k = 0; k < 16; k++) v[k] = s[d * PITCH + 16 * k + (j ^ k)];
            bf16<false>(v);
            c64 *o = dst + row_base + size_t(__brev(lo0 + uint32_t(d)) >> (32 - logm)) * 256 + j;
#pragma unroll
            for (int k = 0; k < 16; k++) st_stream(o + 16 * k, v[k]);
        }
    }
}

template <int TW>
cudaError_t launch_rows_std(bool inverse, const c64 *src, c64 *dst, uint64_t batch, uint32_t n, const c64 *tw_base,
                            cudaStream_t stream)
{
    // TW = 16: the one-exchange kernel above with CFFT_B200_ROWS_STD_ONE_EXCHANGE=1.  Measured (profiles/r2f_ordered_rows_ab.txt): no
    // faster than the two-exchange kernel below (n = 2^16 whole batch 2.89 / 3.02 vs 2.96 / 2.99 TB/s, chunked 3.02 / 3.28 vs
    // 2.98 / 3.30) -- this pass is bound by its 256-byte scatter / gather, M c64 apart, not by the LSU pipe -- so the
    // round-1 kernel stays the default and this one a second implementation the tests compare bit for bit.
    static const bool one_exchange = [] { const char *e = getenv("CFFT_B200_ROWS_STD_ONE_EXCHANGE"); return e && atoi(e) != 0; }();
    if (TW == 16 && one_exchange) {
        const uint32_t m16 = n / 256;
        uint32_t lg = 0;
        while ((1u << lg) < m16) lg++;
        const size_t smem16 = size_t(16) * 257 * sizeof(c64);
        auto fk16 = c64_rows256_std16_kernel<true>;
        auto ik16 = c64_rows256_std16_kernel<false>;
        cudaError_t e = allow_smem(fk16, smem16);
        if (e == cudaSuccess) e = allow_smem(ik16, smem16);
        if (e != cudaSuccess) return e;
        const uint64_t ctas16 = batch * (m16 / 16);
        if (inverse) ik16<<<unsigned(ctas16), 256, smem16, stream>>>(src, dst, n, lg, tw_base);
        else fk16<<<unsigned(ctas16), 256, smem16, stream>>>(src, dst, n, lg, tw_base);
        count_launch();
        return cudaGetLastError();
    }
    const uint32_t m = n / 256;
    uint32_t logm = 0;
    while ((1u << logm) < m) logm++;
    const size_t smem = size_t(TW) * 257 * sizeof(c64);
    auto fk = c64_rows256_std_kernel<TW, true>;
    auto ik = c64_rows256_std_kernel<TW, false>;
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
    const uint64_t ctas = batch * (m / TW);
    if (inverse) ik<<<unsigned(ctas), 16 * TW, smem, stream>>>(src, dst, n, logm, tw_base);
    else fk<<<unsigned(ctas), 16 * TW, smem, stream>>>(src, dst, n, logm, tw_base);
    count_launch();
    return cudaGetLastError();
}

} // namespace

// Stream-ordered workspace for the out-of-place (ordered) path and the composed fused-product path: one private
// pool per device with a bounded release threshold; every user allocates at most kWorkspaceChunkBytes per call.
cudaError_t workspace_pool(int device, cudaMemPool_t *out)
{
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaError_t e = cudaMemPoolCreate(&pools[device], &props);
        if (e != cudaSuccess) return e;
        // keep at most kWorkspaceChunkBytes between calls (steady-state calls of the chunked paths never touch the
        // OS allocator); anything above that goes back to the driver at the next synchronisation point
        uint64_t keep = kWorkspaceChunkBytes;
        e = cudaMemPoolSetAttribute(pools[device], cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
    }
    *out = pools[device];
    return cudaSuccess;
}

// Does a plan qualify?  unordered, base (Dif16, 256), n >= 256.
//   n <= 8192 : one kernel (levels + base FFT fused, one HBM round trip)
//   n >  8192 : column passes for the levels (c64_column.cu), then the base FFTs as rows of 256
bool fast_b256_supported(uint64_t n, int base_algo, uint64_t base_n)
{
    return base_algo == 6 /* Dif16 */ && base_n == 256 && n >= 256 && n <= (uint64_t{1} << 26);
}

// out[r] = inv( sum_k fwd(a[r][k]) (.) b[r * b_row_stride + k n ..] ), r < batch, k < kterms.
//   * plans of the (Dif16, 256) family with n <= 4096: one launch of c64_fwd_mul_inv_kernel;
//   * every other c64 plan: the same arithmetic composed from the plan's own kernels through a stream-ordered
//     workspace (strided copy of term k -> fwd -> point-wise into out, then inv on out), batch in chunks of
//     <= 256 MiB.  Same bits either way.
bool fused_mul_kernel_available(const cfft_plan *plan)
{
    // n = 8192: one kernel per term (512 threads x 128 registers leave no room for a running sum), chained through `out`
    return plan->d_fast_tw[0] && plan->n >= 256 && plan->n <= 8192 &&
           (plan->fast_variant == 1 || plan->fast_variant == 2 || plan->fast_variant == 4);
}

static cudaError_t dispatch_fused_mul(const cfft_plan *plan, const c64 *a, const c64 *b, c64 *out, uint64_t batch, uint32_t kt,
                                      uint64_t b_row_stride, cudaStream_t stream, uint32_t chain)
{
    FastTables tf, ti;
    const c64 *bf = plan->d_fast_tw[0], *bi = plan->d_fast_tw[1];
    tf.top1 = plan->fast_levels.size() > 0 ? bf + plan->fast_levels[0].off : bf;
    tf.top2 = plan->fast_levels.size() > 1 ? bf + plan->fast_levels[1].off : bf;
    tf.base = bf + plan->fast_base_off;
    ti.top1 = plan->fast_levels.size() > 0 ? bi + plan->fast_levels[0].off : bi;
    ti.top2 = plan->fast_levels.size() > 1 ? bi + plan->fast_levels[1].off : bi;
    ti.base = bi + plan->fast_base_off;
    switch (plan->n) {
    case 256: return launch_fused_mul<256, 1, 1>(a, b, out, batch, kt, b_row_stride, tf, ti, stream, chain);
    case 512: return launch_fused_mul<512, 2, 1>(a, b, out, batch, kt, b_row_stride, tf, ti, stream, chain);
    case 1024: return launch_fused_mul<1024, 4, 1>(a, b, out, batch, kt, b_row_stride, tf, ti, stream, chain);
    case 2048: return launch_fused_mul<2048, 8, 1>(a, b, out, batch, kt, b_row_stride, tf, ti, stream, chain);
    case 4096: return launch_fused_mul<4096, 8, 2>(a, b, out, batch, kt, b_row_stride, tf, ti, stream, chain);
    case 8192: return launch_fused_mul<8192, 8, 4, false>(a, b, out, batch, kt, b_row_stride, tf, ti, stream, chain);
    default: return cudaErrorInvalidValue;
    }
}

// acc[r] <- [acc[r] +] fwd(a[r * a_row_terms * n ..]) (.) b[r * b_row_stride ..], all in the Fourier domain (no inverse): the
// building block of loops that produce their terms one at a time, or that feed several accumulators from the same input.
// Plans with the fused kernel: one launch (a read once, acc read + written once); others: copy -> fwd -> point-wise.
cudaError_t launch_c64_fwd_mul_add(const cfft_plan *plan, const double2 *a, uint64_t a_row_terms, const double2 *b,
                                   uint64_t b_row_stride, double2 *acc, bool accumulate, uint64_t batch, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    const bool force_composed = getenv("CFFT_B200_FUSED_MUL_COMPOSED") != nullptr;
    if (fused_mul_kernel_available(plan) && !force_composed && a_row_terms <= 0xFFFFFFFFull)
        return dispatch_fused_mul(plan, a, b, acc, batch, uint32_t(a_row_terms), b_row_stride, stream, 16u | (accumulate ? 8u : 0u));
    const uint64_t n = plan->n, row_bytes = n * sizeof(c64);
    uint64_t chunk_rows = (uint64_t{256} << 20) / row_bytes;
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > batch) chunk_rows = batch;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(plan->device, &pool);
    if (e != cudaSuccess) return e;
    c64 *ws = nullptr;
    e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), chunk_rows * row_bytes, pool, stream);
    if (e != cudaSuccess) return e;
    for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
        const uint64_t rows = batch - r0 < chunk_rows ? batch - r0 : chunk_rows;
        if (a_row_terms == 1) e = cudaMemcpyAsync(ws, a + r0 * n, rows * row_bytes, cudaMemcpyDeviceToDevice, stream);
        else e = cudaMemcpy2DAsync(ws, row_bytes, a + r0 * a_row_terms * n, a_row_terms * row_bytes, row_bytes, rows,
                                   cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = launch_c64(plan, false, ws, rows, stream);
        if (e != cudaSuccess) break;
        uint64_t blocks = (rows * n + 255) / 256;
        if (blocks > 148ull * 32) blocks = 148ull * 32;
        if (accumulate) c64_pointwise_rows_kernel<true><<<unsigned(blocks), 256, 0, stream>>>(acc + r0 * n, ws, b + r0 * b_row_stride, uint32_t(n), rows, b_row_stride);
        else c64_pointwise_rows_kernel<false><<<unsigned(blocks), 256, 0, stream>>>(acc + r0 * n, ws, b + r0 * b_row_stride, uint32_t(n), rows, b_row_stride);
        count_launch();
        e = cudaGetLastError();
    }
    const cudaError_t e2 = cudaFreeAsync(ws, stream);
    return e != cudaSuccess ? e : e2;
}

cudaError_t launch_c64_fwd_mul_inv(const cfft_plan *plan, const double2 *a, uint64_t kterms, const double2 *b,
                                   uint64_t b_row_stride, double2 *out, uint64_t batch, cudaStream_t stream)
{
    if (batch == 0 || kterms == 0) return cudaSuccess;
    const bool force_composed = getenv("CFFT_B200_FUSED_MUL_COMPOSED") != nullptr; // testing hook, read per call
    if (fused_mul_kernel_available(plan) && !force_composed && kterms <= 0xFFFFFFFFull)
        return dispatch_fused_mul(plan, a, b, out, batch, uint32_t(kterms), b_row_stride, stream, 0);
    const uint64_t n = plan->n, row_bytes = n * sizeof(c64);
    uint64_t chunk_rows = (uint64_t{256} << 20) / row_bytes;
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > batch) chunk_rows = batch;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(plan->device, &pool);
    if (e != cudaSuccess) return e;
    c64 *ws = nullptr;
    e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), chunk_rows * row_bytes, pool, stream);
    if (e != cudaSuccess) return e;
    for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
        const uint64_t rows = batch - r0 < chunk_rows ? batch - r0 : chunk_rows;
        c64 *o = out + r0 * n;
        const c64 *bb = b + r0 * b_row_stride;
        for (uint64_t k = 0; k < kterms && e == cudaSuccess; k++) {
            if (kterms == 1) e = cudaMemcpyAsync(ws, a + r0 * n, rows * row_bytes, cudaMemcpyDeviceToDevice, stream);
            else e = cudaMemcpy2DAsync(ws, row_bytes, a + (r0 * kterms + k) * n, kterms * row_bytes, row_bytes, rows,
                                       cudaMemcpyDeviceToDevice, stream);
            if (e == cudaSuccess) e = launch_c64(plan, false, ws, rows, stream);
            if (e != cudaSuccess) break;
            uint64_t blocks = (rows * n + 255) / 256;
            if (blocks > 148ull * 32) blocks = 148ull * 32;
            if (k == 0) c64_pointwise_rows_kernel<false><<<unsigned(blocks), 256, 0, stream>>>(o, ws, bb, uint32_t(n), rows, b_row_stride);
            else c64_pointwise_rows_kernel<true><<<unsigned(blocks), 256, 0, stream>>>(o, ws, bb + k * n, uint32_t(n), rows, b_row_stride);
            count_launch();
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = launch_c64(plan, true, o, rows, stream);
    }
    const cudaError_t e2 = cudaFreeAsync(ws, stream);
    return e != cudaSuccess ? e : e2;
}

// rows row_stride >= n elements apart, for the plans served by ONE fused kernel (fast_variant 1 and 5): the kernel's row
// accessor takes the stride, nothing else changes
bool fast_b256_strided_available(const cfft_plan *plan)
{
    return (plan->fast_variant == 1 || plan->fast_variant == 5) && plan->n >= 256 && plan->n <= 8192;
}
cudaError_t launch_c64_fast_b256_strided(const cfft_plan *plan, bool inverse, double2 *data, uint64_t row_stride, uint64_t batch,
                                         cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    if (!fast_b256_strided_available(plan) || row_stride < plan->n) return cudaErrorInvalidValue;
    const FastTables tb = fast_tables(plan, inverse ? 1 : 0);
    if (plan->fast_variant == 5) {
        switch (plan->n) {
        case 2048: return launch_cfg<2048, 8, 1, true>(inverse, data, batch, tb, stream, row_stride);
        case 4096: return launch_cfg<4096, 8, 2, true>(inverse, data, batch, tb, stream, row_stride);
        case 8192: return launch_cfg<8192, 8, 4, true>(inverse, data, batch, tb, stream, row_stride);
        default: return cudaErrorInvalidValue;
        }
    }
    switch (plan->n) {
    case 256: return launch_cfg<256, 1, 1>(inverse, data, batch, tb, stream, row_stride);
    case 512: return launch_cfg<512, 2, 1>(inverse, data, batch, tb, stream, row_stride);
    case 1024: return launch_cfg<1024, 4, 1>(inverse, data, batch, tb, stream, row_stride);
    case 2048: return launch_cfg<2048, 8, 1>(inverse, data, batch, tb, stream, row_stride);
    case 4096: return launch_cfg<4096, 8, 2>(inverse, data, batch, tb, stream, row_stride);
    case 8192: return launch_cfg<8192, 8, 4>(inverse, data, batch, tb, stream, row_stride);
    default: return cudaErrorInvalidValue;
    }
}

// ---- several outputs per row: out[r][o] = inv( sum_k fwd(a[r][k]) (.) b[r][k][o] ), o < n_out (the GLWE external product) ----
template <int N, int R1, int R2>
static cudaError_t launch_fused_mul2(const cfft_plan *plan, const c64 *a, const c64 *b, c64 *out, uint64_t batch, uint32_t kterms,
                                     uint64_t b_row_stride, cudaStream_t stream)
{
    using Cfg = FastCfg<N>;
    const size_t smem = size_t(Cfg::ROWS) * N * sizeof(c64);
    auto k = c64_fwd_mul_inv2_kernel<N, R1, R2>;
    cudaError_t e = allow_smem(k, smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t ctas = (batch + Cfg::ROWS - 1) / Cfg::ROWS;
    k<<<unsigned(ctas), Cfg::NT, smem, stream>>>(a, b, out, batch, kterms, b_row_stride, fast_tables(plan, 0), fast_tables(plan, 1),
                                                 N >= 2048 ? uint32_t(sms * 2 * Cfg::ROWS) : 0u);
    count_launch();
    return cudaGetLastError();
}

// one kernel for two outputs: the (Dif16, 256) plans of n = 512, 1024, 2048 (TFHE polynomial sizes 1024 .. 4096)
bool fused_mul2_kernel_available(const cfft_plan *plan)
{
    return fused_mul_kernel_available(plan) && (plan->n == 512 || plan->n == 1024 || plan->n == 2048);
}

cudaError_t launch_c64_fwd_mul_inv_multi(const cfft_plan *plan, const double2 *a, uint64_t kterms, const double2 *b, uint64_t b_row_stride,
                                         uint64_t n_out, double2 *out, uint64_t batch, cudaStream_t stream)
{
    if (batch == 0 || kterms == 0 || n_out == 0) return cudaSuccess;
    if (n_out == 1) return launch_c64_fwd_mul_inv(plan, a, kterms, b, b_row_stride, out, batch, stream);
    const uint64_t n = plan->n;
    const bool force_composed = getenv("CFFT_B200_FUSED_MUL_COMPOSED") != nullptr; // testing hook, read per call
    if (n_out == 2 && fused_mul2_kernel_available(plan) && !force_composed && kterms <= 0xFFFFFFFFull) {
        switch (n) {
        case 512: return launch_fused_mul2<512, 2, 1>(plan, a, b, out, batch, uint32_t(kterms), b_row_stride, stream);
        case 1024: return launch_fused_mul2<1024, 4, 1>(plan, a, b, out, batch, uint32_t(kterms), b_row_stride, stream);
        default: return launch_fused_mul2<2048, 8, 1>(plan, a, b, out, batch, uint32_t(kterms), b_row_stride, stream);
        }
    }
    // every other case: one output at a time through the single-output call (the forward transforms are recomputed per
    // output), with the operand rows of that output gathered into, and the results scattered from, the stream-ordered workspace
    const uint64_t row_bytes = n * sizeof(c64);
    const bool shared_b = b_row_stride == 0;
    uint64_t chunk_rows = (uint64_t{128} << 20) / (row_bytes * (shared_b ? 1 : kterms + 1));
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > batch) chunk_rows = batch;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(plan->device, &pool);
    if (e != cudaSuccess) return e;
    c64 *ws_out = nullptr, *ws_b = nullptr;
    if ((e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws_out), chunk_rows * row_bytes, pool, stream)) != cudaSuccess) return e;
    e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws_b), (shared_b ? 1 : chunk_rows) * kterms * row_bytes, pool, stream);
    if (e != cudaSuccess) {
        cudaFreeAsync(ws_out, stream);
        return e;
    }
    for (uint64_t o = 0; o < n_out && e == cudaSuccess; o++) {
        if (shared_b) e = cudaMemcpy2DAsync(ws_b, row_bytes, b + o * n, n_out * row_bytes, row_bytes, kterms, cudaMemcpyDeviceToDevice, stream);
        for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
            const uint64_t rows = batch - r0 < chunk_rows ? batch - r0 : chunk_rows;
            if (!shared_b) {
                if (b_row_stride == kterms * n_out * n) {
                    e = cudaMemcpy2DAsync(ws_b, row_bytes, b + r0 * b_row_stride + o * n, n_out * row_bytes, row_bytes, rows * kterms, cudaMemcpyDeviceToDevice, stream);
                } else {
                    for (uint64_t r = 0; r < rows && e == cudaSuccess; r++)
                        e = cudaMemcpy2DAsync(ws_b + r * kterms * n, row_bytes, b + (r0 + r) * b_row_stride + o * n, n_out * row_bytes, row_bytes, kterms,
                                              cudaMemcpyDeviceToDevice, stream);
                }
            }
            if (e == cudaSuccess) e = launch_c64_fwd_mul_inv(plan, a + r0 * kterms * n, kterms, ws_b, shared_b ? 0 : kterms * n, ws_out, rows, stream);
            if (e == cudaSuccess)
                e = cudaMemcpy2DAsync(out + (r0 * n_out + o) * n, n_out * row_bytes, ws_out, row_bytes, row_bytes, rows, cudaMemcpyDeviceToDevice, stream);
        }
    }
    const cudaError_t e2 = cudaFreeAsync(ws_out, stream), e3 = cudaFreeAsync(ws_b, stream);
    return e != cudaSuccess ? e : (e2 != cudaSuccess ? e2 : e3);
}

cudaError_t launch_c64_fast_b256(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    const int d = inverse ? 1 : 0;
    const c64 *base = plan->d_fast_tw[d];
    FastTables tb;
    tb.top1 = plan->fast_levels.size() > 0 ? base + plan->fast_levels[0].off : base;
    tb.top2 = plan->fast_levels.size() > 1 ? base + plan->fast_levels[1].off : base;
    tb.base = base + plan->fast_base_off;
    if (plan->fast_variant == 2 || plan->fast_variant == 3 || plan->fast_variant == 9) {
        // Multi-pass variants: unordered (2, in place) and ordered (3: the last column pass goes out of
        // place into a workspace and the base-FFT pass writes standard order back).
        //
        // Scheduling (plan->l2_chunk_mb / l2_streams, chosen by the autotuner or forced with
        // CFFT_B200_L2_CHUNK_MB / CFFT_B200_L2_STREAMS): by default each pass covers the whole batch.
        // Otherwise the passes run chunk by chunk so that what one pass wrote is still in the 126 MB L2
        // when the next reads it; chunks alternate over 2-4 auxiliary streams (fork / join on events)
        // because back-to-back small kernels on ONE stream lose more to launch gaps than L2 gains.
        static const long env_mb = [] { const char *e = getenv("CFFT_B200_L2_CHUNK_MB"); return e ? atol(e) : -1; }();
        static const int env_streams = [] { const char *e = getenv("CFFT_B200_L2_STREAMS"); return e ? atoi(e) : -1; }();
        const bool ordered = plan->fast_variant == 3;
        const long mb = env_mb >= 0 ? env_mb : long(plan->l2_chunk_mb);
        int nstreams = env_streams >= 0 ? env_streams : int(plan->l2_streams);
        nstreams = nstreams < 1 ? 1 : (nstreams > 4 ? 4 : nstreams);
        const uint64_t chunk_bytes = uint64_t(mb > 0 ? mb : 1 << 20) << 20;
        uint64_t rows_per_chunk = chunk_bytes / (plan->n * sizeof(c64));
        if (rows_per_chunk < 1) rows_per_chunk = 1;
        const uint32_t n32 = uint32_t(plan->n);
        // variant 9: fewer column passes, then the fused kernel of tail_n points on contiguous blocks
        const bool fused_tail = plan->fast_variant == 9;
        const std::vector<cfft_plan::FastGroup> &groups = fused_tail ? plan->tail_groups : plan->fast_groups;
        const size_t ng = groups.size();
        FastTables tbt = tb;
        if (fused_tail) {
            const size_t l0 = size_t(plan->tail_first_level);
            tbt.top1 = base + plan->fast_levels[l0].off;
            tbt.top2 = l0 + 1 < plan->fast_levels.size() ? base + plan->fast_levels[l0 + 1].off : base;
        }
        auto tail_pass = [&](c64 *d0, uint64_t nrows, cudaStream_t st) -> cudaError_t {
            if (!fused_tail) return launch_cfg<256, 1, 1>(inverse, d0, nrows * (n32 / 256), tb, st);
            const uint64_t blocks = nrows * (n32 / plan->tail_n);
            switch (plan->tail_n) {
            case 512: return launch_cfg<512, 2, 1>(inverse, d0, blocks, tbt, st);
            case 1024: return launch_cfg<1024, 4, 1>(inverse, d0, blocks, tbt, st);
            case 2048: return launch_cfg<2048, 8, 1>(inverse, d0, blocks, tbt, st);
            case 4096: return launch_cfg<4096, 8, 2>(inverse, d0, blocks, tbt, st);
            default: return cudaErrorInvalidValue;
            }
        };
        cudaMemPool_t pool = nullptr;
        if (ordered) {
            cudaError_t e = workspace_pool(plan->device, &pool);
            if (e != cudaSuccess) return e;
        }

        // all passes of `rows` transforms starting at d0, on stream st
        auto process = [&](c64 *d0, uint64_t rows, cudaStream_t st) -> cudaError_t {
            auto group = [&](const cfft_plan::FastGroup &g, const c64 *src, c64 *dst) {
                const double2 *tw[3] = {base, base, base};
                for (int i = 0; i < 3; i++)
                    if (g.radices[i] > 1) tw[i] = base + plan->fast_levels[size_t(g.first_level + i)].off;
                return launch_c64_column_group(inverse, src, dst, rows, n32, g.span0, g.radices, tw, st);
            };
            cudaError_t e = cudaSuccess;
            if (!ordered) {
                if (!inverse) {
                    for (size_t i = 0; i < ng && e == cudaSuccess; i++) e = group(groups[i], d0, d0);
                    if (e == cudaSuccess) e = tail_pass(d0, rows, st);
                } else {
                    e = tail_pass(d0, rows, st);
                    for (size_t i = ng; i-- > 0 && e == cudaSuccess;) e = group(groups[i], d0, d0);
                }
                return e;
            }
            // the out-of-place workspace is bounded (kWorkspaceChunkBytes): a larger call runs slice by slice on the
            // same stream, so the pool never holds more than that on behalf of one call
            uint64_t ws_rows = kWorkspaceChunkBytes / (plan->n * sizeof(c64));
            if (ws_rows < 1) ws_rows = 1;
            if (ws_rows > rows) ws_rows = rows;
            c64 *ws = nullptr;
            e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), ws_rows * plan->n * sizeof(c64), pool, st);
            if (e != cudaSuccess) return e;
            for (uint64_t r0 = 0; r0 < rows && e == cudaSuccess; r0 += ws_rows) {
                const uint64_t nr = rows - r0 < ws_rows ? rows - r0 : ws_rows;
                c64 *dd = d0 + r0 * plan->n;
                auto group_n = [&](const cfft_plan::FastGroup &g, const c64 *src, c64 *dst) {
                    const double2 *tw[3] = {base, base, base};
                    for (int i = 0; i < 3; i++)
                        if (g.radices[i] > 1) tw[i] = base + plan->fast_levels[size_t(g.first_level + i)].off;
                    return launch_c64_column_group(inverse, src, dst, nr, n32, g.span0, g.radices, tw, st);
                };
                auto rows_pass = [&](const c64 *src, c64 *dst) {
                    return n32 / 256 >= 16 ? launch_rows_std<16>(inverse, src, dst, nr, n32, tb.base, st)
                                           : launch_rows_std<8>(inverse, src, dst, nr, n32, tb.base, st);
                };
                if (!inverse) {
                    for (size_t i = 0; i < ng && e == cudaSuccess; i++)
                        e = group_n(plan->fast_groups[i], dd, i + 1 == ng ? ws : dd);
                    if (e == cudaSuccess) e = rows_pass(ws, dd);
                } else {
                    e = rows_pass(dd, ws);
                    for (size_t i = ng; i-- > 0 && e == cudaSuccess;)
                        e = group_n(plan->fast_groups[i], i + 1 == ng ? ws : dd, dd);
                }
            }
            const cudaError_t e2 = cudaFreeAsync(ws, st);
            return e != cudaSuccess ? e : e2;
        };

        if (rows_per_chunk >= batch) return process(data, batch, stream);

        // auxiliary streams / events of the fork-join: one set per (host thread, device), created once and kept
        struct Aux { bool ready = false; cudaStream_t st[4] = {}; cudaEvent_t done[4] = {}; cudaEvent_t start = nullptr; };
        static thread_local Aux aux_by_device[64];
        const bool fork = nstreams > 1;
        if (fork && (plan->device < 0 || plan->device >= 64)) return cudaErrorInvalidDevice;
        Aux &aux = aux_by_device[fork ? plan->device : 0];
        if (fork && !aux.ready) {
            cudaError_t ce = cudaSuccess;
            for (int i = 0; i < 4 && ce == cudaSuccess; i++) {
                ce = cudaStreamCreateWithFlags(&aux.st[i], cudaStreamNonBlocking);
                if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&aux.done[i], cudaEventDisableTiming);
            }
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&aux.start, cudaEventDisableTiming);
            if (ce != cudaSuccess) { // nothing half-built is kept
                for (int i = 0; i < 4; i++) {
                    if (aux.st[i]) cudaStreamDestroy(aux.st[i]);
                    if (aux.done[i]) cudaEventDestroy(aux.done[i]);
                }
                if (aux.start) cudaEventDestroy(aux.start);
                aux = Aux{};
                return ce;
            }
            aux.ready = true;
        }
        cudaError_t e = cudaSuccess;
        if (fork) {
            e = cudaEventRecord(aux.start, stream);
            for (int i = 0; i < nstreams && e == cudaSuccess; i++) e = cudaStreamWaitEvent(aux.st[i], aux.start, 0);
            if (e != cudaSuccess) return e;
        }
        uint64_t chunk_index = 0;
        for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += rows_per_chunk, chunk_index++) {
            const uint64_t rows = (batch - r0 < rows_per_chunk) ? batch - r0 : rows_per_chunk;
            e = process(data + r0 * plan->n, rows, fork ? aux.st[chunk_index % uint64_t(nstreams)] : stream);
        }
        if (fork) { // always join, even after a failed launch: the caller's stream must stay ordered after the side streams
            for (int i = 0; i < nstreams; i++) {
                cudaError_t je = cudaEventRecord(aux.done[i], aux.st[i]);
                if (je == cudaSuccess) je = cudaStreamWaitEvent(stream, aux.done[i], 0);
                if (e == cudaSuccess) e = je;
            }
        }
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    if (plan->fast_variant == 8) { // n = 2^14 .. 2^16, one persistent kernel for both passes
        static const int env_lag = [] { const char *e = getenv("CFFT_B200_TWOPASS_LAG"); return e ? atoi(e) : 0; }();
        if (plan->fast_groups.size() != 1) return cudaErrorInvalidValue;
        const cfft_plan::FastGroup &g = plan->fast_groups[0];
        const double2 *tw[3] = {base, base, base};
        for (int i = 0; i < 3; i++)
            if (g.radices[i] > 1) tw[i] = base + plan->fast_levels[size_t(g.first_level + i)].off;
        return launch_c64_twopass(inverse, data, batch, uint32_t(plan->n), g.radices, tw, tb.base, uint32_t(env_lag > 0 ? env_lag : 0),
                                  plan->device, stream);
    }
    if (plan->fast_variant == 5) { // standard-order in / out, one kernel (ordered plans 2^11 <= n <= 2^13)
        switch (plan->n) {
        case 2048: return launch_cfg<2048, 8, 1, true>(inverse, data, batch, tb, stream);
        case 4096: return launch_cfg<4096, 8, 2, true>(inverse, data, batch, tb, stream);
        case 8192: return launch_cfg<8192, 8, 4, true>(inverse, data, batch, tb, stream);
        default: return cudaErrorInvalidValue;
        }
    }
    if (plan->fast_variant == 4) {
        if (plan->n == 8192) return launch_cluster<8192, 2, 4>(inverse, data, batch, tb, stream);
        if (plan->n == 16384) return launch_cluster<16384, 4, 8>(inverse, data, batch, tb, stream);
        return cudaErrorInvalidValue;
    }
    switch (plan->n) {
    case 256: return launch_cfg<256, 1, 1>(inverse, data, batch, tb, stream);
    case 512: return launch_cfg<512, 2, 1>(inverse, data, batch, tb, stream);
    case 1024: return launch_cfg<1024, 4, 1>(inverse, data, batch, tb, stream);
    case 2048: return launch_cfg<2048, 8, 1>(inverse, data, batch, tb, stream);
    case 4096: return launch_cfg<4096, 8, 2>(inverse, data, batch, tb, stream);
    case 8192: return launch_cfg<8192, 8, 4>(inverse, data, batch, tb, stream);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace cfft

// c64_tile.cu -- the "exact" c64 kernels: execute the reference's stage schedule for ANY
// (base_algo, base_n) plan, butterfly for butterfly, so the output is bit-identical to
// concrete-fft's for the same plan (and lands in the same permuted order).
//
//   c64_tile_kernel   one CTA owns a tile of <= 4096 contiguous c64 held in shared memory
//                     (two ping-pong buffers) and runs every stage whose span fits the tile:
//                     unordered levels  src/unordered.rs:222-293 (fwd_process_x*, inv_process_x*)
//                     Stockham stages   src/dif{2,4,8,16}.rs / src/dit{2,4,8,16}.rs (_generic + _end)
//   c64_global_stage  unordered levels whose span exceeds a tile (N >= 8192), one in-place pass
//                     over HBM per level -- the reference recursion's top levels.
//   monomial / permute kernels: src/unordered.rs:844-900 and :942-1036.
//
// HBM traffic of the tile kernel is the algorithmic minimum (read N, write N c64 per transform);
// twiddles come from the plan's tables through L1/L2.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "c64_math.cuh"
#include "plan.h"

namespace cfft {

namespace {

constexpr int kThreads = 256;

// Shared-memory index swizzle: XOR the low three bits (the 16-byte bank group) with bits 3..5 and
// 6..8.  Stage patterns touch either consecutive elements or elements R = 2..16 apart across lanes;
// both become conflict-free for the eight lanes of a 128-bit phase.
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u) ^ ((i >> 6) & 7u); }

__device__ __forceinline__ uint32_t brev_small(uint32_t k, int bits)
{
    return __brev(k) >> (32 - bits);
}

// ---- in-tile stages -------------------------------------------------------------------------

// unordered level on blocks of n_cur elements, in place.
// fwd: v = DFT_r(z[p + m k]); z[p + m bitrev(k)] = w_k v_k          src/unordered.rs:244-278
// inv: v_k = w_k z[p + m bitrev(k)]; z[p + m k] = DFT_r^-1(v)_k     src/unordered.rs:254-293
template <int R, bool FWD>
__device__ __forceinline__ void stage_top(c64 *buf, uint32_t valid, uint32_t n_cur, const c64 *__restrict__ w)
{
    constexpr int RB = (R == 2) ? 1 : (R == 4 ? 2 : 3);
    const uint32_t m = n_cur / R;
    for (uint32_t b = threadIdx.x; b < valid / R; b += kThreads) {
        const uint32_t blk = b / m, p = b - blk * m;
        const uint32_t z = blk * n_cur + p;
        const c64 *wp = w + (R - 1) * p;
        c64 v[R];
        if (FWD) {
#pragma unroll
            for (int k = 0; k < R; k++) v[k] = buf[swz(z + m * k)];
            bfR<R, true>(v);
            buf[swz(z)] = v[0];
#pragma unroll
            for (int k = 1; k < R; k++) buf[swz(z + m * brev_small(k, RB))] = cmul(wp[k - 1], v[k]);
        } else {
            v[0] = buf[swz(z)];
#pragma unroll
            for (int k = 1; k < R; k++) v[k] = cmul(wp[k - 1], buf[swz(z + m * brev_small(k, RB))]);
            bfR<R, false>(v);
#pragma unroll
            for (int k = 0; k < R; k++) buf[swz(z + m * k)] = v[k];
        }
    }
}

// Stockham DIF stage on blocks of base_n: x -> y.
// y[q + s(R p + k)] = w[R p s + k] * DFT_R(x[q + s(p + m k)])_k
template <int R, bool FWD>
__device__ __forceinline__ void stage_core_dif(const c64 *x, c64 *y, uint32_t valid, uint32_t base_n, uint32_t s,
                                               const c64 *__restrict__ w)
{
    const uint32_t per_blk = base_n / R, m = per_blk / s;
    for (uint32_t b = threadIdx.x; b < valid / R; b += kThreads) {
        const uint32_t blk = b / per_blk, rem = b - blk * per_blk;
        const uint32_t p = rem / s, q = rem - p * s;
        const uint32_t xi = blk * base_n + q + s * p;
        const uint32_t yo = blk * base_n + q + s * R * p;
        const c64 *wp = w + R * p * s;
        c64 v[R];
#pragma unroll
        for (int k = 0; k < R; k++) v[k] = x[swz(xi + s * m * k)];
        bfR<R, FWD>(v);
        y[swz(yo)] = v[0];
#pragma unroll
        for (int k = 1; k < R; k++) y[swz(yo + s * k)] = cmul(wp[k], v[k]);
    }
}

// Stockham DIT stage on blocks of base_n: y -> x.
// x[q + s(p + m k)] = DFT_R(w[R p s + k] * y[q + s(R p + k)])_k
template <int R, bool FWD>
__device__ __forceinline__ void stage_core_dit(const c64 *y, c64 *x, uint32_t valid, uint32_t base_n, uint32_t s,
                                               const c64 *__restrict__ w)
{
    const uint32_t per_blk = base_n / R, m = per_blk / s;
    for (uint32_t b = threadIdx.x; b < valid / R; b += kThreads) {
        const uint32_t blk = b / per_blk, rem = b - blk * per_blk;
        const uint32_t p = rem / s, q = rem - p * s;
        const uint32_t yi = blk * base_n + q + s * R * p;
        const uint32_t xo = blk * base_n + q + s * p;
        const c64 *wp = w + R * p * s;
        c64 v[R];
        v[0] = y[swz(yi)];
#pragma unroll
        for (int k = 1; k < R; k++) v[k] = cmul(wp[k], y[swz(yi + s * k)]);
        bfR<R, FWD>(v);
#pragma unroll
        for (int k = 0; k < R; k++) x[swz(xo + s * m * k)] = v[k];
    }
}

// terminal twiddle-free pass, in place
template <int R, bool FWD>
__device__ __forceinline__ void stage_end(c64 *buf, uint32_t valid, uint32_t base_n)
{
    const uint32_t part = base_n / R;
    for (uint32_t b = threadIdx.x; b < valid / R; b += kThreads) {
        const uint32_t blk = b / part, j = b - blk * part;
        const uint32_t z = blk * base_n + j;
        c64 v[R];
#pragma unroll
        for (int k = 0; k < R; k++) v[k] = buf[swz(z + part * k)];
        bfR<R, FWD>(v);
#pragma unroll
        for (int k = 0; k < R; k++) buf[swz(z + part * k)] = v[k];
    }
}

template <bool FWD>
__global__ void __launch_bounds__(kThreads)
c64_tile_kernel(c64 *__restrict__ data, uint64_t total, uint32_t tile, uint32_t base_n, StageProgram prog,
                const c64 *__restrict__ tw)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c64 *cur = reinterpret_cast<c64 *>(smem_raw);
    c64 *oth = cur + ((tile + 7u) & ~7u); // swz() permutes inside aligned groups of 8 elements

    const uint64_t start = uint64_t(blockIdx.x) * tile;
    const uint32_t valid = (total - start < tile) ? uint32_t(total - start) : tile;
    c64 *g = data + start;

    for (uint32_t i = threadIdx.x; i < valid; i += kThreads) cur[swz(i)] = g[i];
    __syncthreads();

    for (int si = 0; si < prog.count; si++) {
        const Stage st = prog.st[si];
        const c64 *w = tw + st.tw_off;
        if (st.kind == ST_TOP) {
            if (st.radix == 8) stage_top<8, FWD>(cur, valid, st.span, w);
            else if (st.radix == 4) stage_top<4, FWD>(cur, valid, st.span, w);
            else stage_top<2, FWD>(cur, valid, st.span, w);
        } else if (st.kind == ST_END) {
            if (st.radix == 16) stage_end<16, FWD>(cur, valid, base_n);
            else if (st.radix == 8) stage_end<8, FWD>(cur, valid, base_n);
            else if (st.radix == 4) stage_end<4, FWD>(cur, valid, base_n);
            else stage_end<2, FWD>(cur, valid, base_n);
        } else {
            if (st.kind == ST_CORE_DIF) {
                if (st.radix == 16) stage_core_dif<16, FWD>(cur, oth, valid, base_n, st.span, w);
                else if (st.radix == 8) stage_core_dif<8, FWD>(cur, oth, valid, base_n, st.span, w);
                else if (st.radix == 4) stage_core_dif<4, FWD>(cur, oth, valid, base_n, st.span, w);
                else stage_core_dif<2, FWD>(cur, oth, valid, base_n, st.span, w);
            } else {
                if (st.radix == 16) stage_core_dit<16, FWD>(cur, oth, valid, base_n, st.span, w);
                else if (st.radix == 8) stage_core_dit<8, FWD>(cur, oth, valid, base_n, st.span, w);
                else if (st.radix == 4) stage_core_dit<4, FWD>(cur, oth, valid, base_n, st.span, w);
                else stage_core_dit<2, FWD>(cur, oth, valid, base_n, st.span, w);
            }
            c64 *t = cur; cur = oth; oth = t;
        }
        __syncthreads();
    }

    for (uint32_t i = threadIdx.x; i < valid; i += kThreads) g[i] = cur[swz(i)];
}

// ---- unordered level through HBM (span > tile) ----------------------------------------------
template <int R, bool FWD>
__global__ void __launch_bounds__(kThreads)
c64_global_stage(c64 *__restrict__ data, uint64_t total, uint32_t n_cur, const c64 *__restrict__ w)
{
    constexpr int RB = (R == 2) ? 1 : (R == 4 ? 2 : 3);
    const uint32_t m = n_cur / R;
    const uint64_t nb = total / R;
    for (uint64_t b = uint64_t(blockIdx.x) * kThreads + threadIdx.x; b < nb; b += uint64_t(gridDim.x) * kThreads) {
        const uint64_t blk = b / m;
        const uint32_t p = uint32_t(b - blk * m);
        c64 *z = data + blk * n_cur + p;
        const c64 *wp = w + (R - 1) * p;
        c64 v[R];
        if (FWD) {
#pragma unroll
            for (int k = 0; k < R; k++) v[k] = z[size_t(m) * k];
            bfR<R, true>(v);
            z[0] = v[0];
#pragma unroll
            for (int k = 1; k < R; k++) z[size_t(m) * brev_small(k, RB)] = cmul(wp[k - 1], v[k]);
        } else {
            v[0] = z[0];
#pragma unroll
            for (int k = 1; k < R; k++) v[k] = cmul(wp[k - 1], z[size_t(m) * brev_small(k, RB)]);
            bfR<R, false>(v);
#pragma unroll
            for (int k = 0; k < R; k++) z[size_t(m) * k] = v[k];
        }
    }
}

// ---- fwd_monomial: buf[i] = tw[(idx(i) * degree) & (n - 1)]   src/unordered.rs:871-891 ---------
__global__ void monomial_kernel(c64 *__restrict__ buf, const c64 *__restrict__ tw, uint32_t n, uint32_t nbits,
                                uint32_t base_nbits, uint32_t degree)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // idx(i) = bit_rev_twice_inv(i) (src/unordered.rs:1054-1059); the n == base_n and
    // n == 2 base_n special cases of :872-885 are this same map written out.
    const uint32_t mask = (1u << base_nbits) - 1;
    const uint32_t low = base_nbits ? (__brev(i & mask) >> (32 - base_nbits)) : 0;
    const uint32_t t = (i & ~mask) | low;
    const uint32_t idx = nbits ? (__brev(t) >> (32 - nbits)) : 0;
    buf[i] = tw[(uint64_t(idx) * degree) & (n - 1)];
}

// ---- standard-order gather / scatter   src/unordered.rs:967-969, 1019-1022 --------------------
template <bool TO_STANDARD>
__global__ void permute_kernel(const c64 *__restrict__ src, c64 *__restrict__ dst, uint64_t total, uint32_t n,
                               uint32_t nbits, uint32_t base_nbits)
{
    for (uint64_t e = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t row = e / n;
        const uint32_t i = uint32_t(e - row * n);
        // bit_rev_twice(i), src/unordered.rs:1046-1051
        const uint32_t irev = nbits ? (__brev(i) >> (32 - nbits)) : 0;
        const uint32_t mask = (1u << base_nbits) - 1;
        const uint32_t low = base_nbits ? (__brev(irev & mask) >> (32 - base_nbits)) : 0;
        const uint32_t pos = (irev & ~mask) | low;
        if (TO_STANDARD) dst[e] = src[row * n + pos];
        else dst[row * n + pos] = src[e];
    }
}

// ---- element-wise Fourier-domain products (num_complex `*`, `+`: no FMA) ---------------------------
template <bool ACC>
__global__ void __launch_bounds__(256)
c64_pointwise_kernel(c64 *out, const c64 *a, const c64 *__restrict__ b, uint64_t len) // out may be a
{
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < len; i += uint64_t(gridDim.x) * blockDim.x) {
        const c64 x = a[i], y = b[i];
        c64 r = mk(__dsub_rn(__dmul_rn(x.x, y.x), __dmul_rn(x.y, y.y)), __dadd_rn(__dmul_rn(x.x, y.y), __dmul_rn(x.y, y.x)));
        if (ACC) r = cadd(out[i], r);
        out[i] = r;
    }
}

template <bool FWD>
cudaError_t launch_global_stage(const Stage &st, c64 *data, uint64_t total, const c64 *tw, cudaStream_t stream)
{
    const uint64_t nb = total / st.radix;
    uint64_t blocks = (nb + kThreads - 1) / kThreads;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    const dim3 grid(static_cast<unsigned>(blocks));
    const c64 *w = tw + st.tw_off;
    if (st.radix == 8) c64_global_stage<8, FWD><<<grid, kThreads, 0, stream>>>(data, total, st.span, w);
    else if (st.radix == 4) c64_global_stage<4, FWD><<<grid, kThreads, 0, stream>>>(data, total, st.span, w);
    else c64_global_stage<2, FWD><<<grid, kThreads, 0, stream>>>(data, total, st.span, w);
    count_launch();
    return cudaGetLastError();
}

template <bool FWD>
cudaError_t launch_tile(const StageProgram &prog, c64 *data, uint64_t total, uint32_t tile, uint32_t base_n,
                        const c64 *tw, cudaStream_t stream)
{
    const size_t smem = size_t((tile + 7u) & ~7u) * sizeof(c64) * 2;
    static thread_local int configured_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device != dev) {
        cudaError_t e = cudaFuncSetAttribute(c64_tile_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             int(kTileMax * sizeof(c64) * 2));
        if (e != cudaSuccess) return e;
        configured_device = dev;
    }
    const uint64_t tiles = (total + tile - 1) / tile;
    c64_tile_kernel<FWD><<<dim3(static_cast<unsigned>(tiles)), kThreads, smem, stream>>>(data, total, tile, base_n,
                                                                                         prog, tw);
    count_launch();
    return cudaGetLastError();
}

} // namespace

cudaError_t launch_c64_exact(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t stream)
{
    const uint64_t n = plan->n;
    if (n <= 1 || batch == 0) return cudaSuccess; // src/ordered.rs:210-212: n == 1 is a no-op
    const uint64_t total = n * batch;
    const StageProgram &full = plan->prog[inverse ? 1 : 0];
    const c64 *tw = plan->d_tw[inverse ? 1 : 0];

    // tile = whole transforms when they fit, else a 4096-element sub-block of one transform
    uint32_t tile;
    const bool regs = plan->exact_regs && !getenv("CFFT_B200_EXACT_TILE"); // register kernel (c64_regs.cu), the default
    if (n >= kTileMax) tile = kTileMax;
    else if (regs) tile = (n <= 1024 && !getenv("CFFT_B200_REGS_TILE2048")) ? 1024 : 2048; // whole transforms; smaller tiles = more CTAs per SM
    else {
        uint64_t rows = (plan->tile_elems ? plan->tile_elems : 2048u) / n;
        if (rows < 1) rows = 1;
        if (rows > batch) rows = batch;
        tile = uint32_t(rows * n);
    }

    StageProgram in_tile;
    in_tile.count = 0;
    for (int i = 0; i < full.count; i++)
        if (!(full.st[i].kind == ST_TOP && full.st[i].span > tile)) in_tile.st[in_tile.count++] = full.st[i];

    cudaError_t e;
    if (regs && (plan->kind == KIND_UNORDERED || !plan->allow_large) && n <= kTileMax && !getenv("CFFT_B200_REGS_NO_SPEC")) {
        // plans with a compile-time schedule (c64_regs.cu): same stages, same tables, index arithmetic folded away
        bool taken = false;
        e = launch_c64_regs_spec(inverse, n, algo_radix(plan->algo), algo_is_dit(plan->algo), plan->kind == KIND_ORDERED ? n : plan->base_n, full, data, total, tw,
                                 plan->d_top_tw[inverse ? 1 : 0], stream, &taken);
        if (e != cudaSuccess || taken) return e;
    }
    if (regs) {
        // Levels wider than the tile (all radix 8: a radix-2 / 4 level spans <= 4 base_n <= 4096) run as
        // column passes over HBM, two levels per pass (c64_column.cu, planar twiddles), outermost pair first.
        std::vector<Stage> wide; // outermost level first
        for (int i = 0; i < full.count; i++)
            if (full.st[i].kind == ST_TOP && full.st[i].span > tile) wide.push_back(full.st[i]);
        if (inverse) std::reverse(wide.begin(), wide.end()); // prog[1] lists them innermost first
        struct Pass { int radices[3]; const double2 *tw[3]; uint32_t span0; };
        std::vector<Pass> passes;
        const double2 *top = plan->d_top_tw[inverse ? 1 : 0];
        for (size_t i = 0; i < wide.size(); i += 2) {
            Pass ps = {{wide[i].radix, 1, 1}, {top + wide[i].tw2, top, top}, wide[i].span};
            if (i + 1 < wide.size()) {
                ps.radices[1] = wide[i + 1].radix;
                ps.tw[1] = top + wide[i + 1].tw2;
            }
            passes.push_back(ps);
        }
        auto column = [&](const Pass &ps) {
            return launch_c64_column_group(inverse, data, data, batch, uint32_t(n), ps.span0, ps.radices, ps.tw, stream);
        };
        if (!inverse) {
            for (const Pass &ps : passes)
                if ((e = column(ps)) != cudaSuccess) return e;
            return launch_c64_regs(false, tile, in_tile, data, total, uint32_t(plan->base_n), tw, top, stream);
        }
        if ((e = launch_c64_regs(true, tile, in_tile, data, total, uint32_t(plan->base_n), tw, top, stream)) != cudaSuccess) return e;
        for (size_t i = passes.size(); i-- > 0;)
            if ((e = column(passes[i])) != cudaSuccess) return e;
        return cudaSuccess;
    }
    if (!inverse) {
        for (int i = 0; i < full.count; i++)
            if (full.st[i].kind == ST_TOP && full.st[i].span > tile)
                if ((e = launch_global_stage<true>(full.st[i], data, total, tw, stream)) != cudaSuccess) return e;
        return launch_tile<true>(in_tile, data, total, tile, uint32_t(plan->base_n), tw, stream);
    }
    if ((e = launch_tile<false>(in_tile, data, total, tile, uint32_t(plan->base_n), tw, stream)) != cudaSuccess) return e;
    for (int i = 0; i < full.count; i++)
        if (full.st[i].kind == ST_TOP && full.st[i].span > tile)
            if ((e = launch_global_stage<false>(full.st[i], data, total, tw, stream)) != cudaSuccess) return e;
    return cudaSuccess;
}

// acc == nullptr: a <- a * b (a is `out`); else acc <- acc + a * b
cudaError_t launch_c64_pointwise(double2 *acc, double2 *a, const double2 *b, uint64_t len, cudaStream_t stream)
{
    if (len == 0) return cudaSuccess;
    uint64_t blocks = (len + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    if (acc) c64_pointwise_kernel<true><<<unsigned(blocks), 256, 0, stream>>>(acc, a, b, len);
    else c64_pointwise_kernel<false><<<unsigned(blocks), 256, 0, stream>>>(a, a, b, len);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_monomial(const cfft_plan *plan, uint64_t degree, double2 *data, cudaStream_t stream)
{
    const uint32_t n = uint32_t(plan->n);
    const unsigned threads = 256, blocks = (n + threads - 1) / threads;
    monomial_kernel<<<blocks, threads, 0, stream>>>(data, plan->d_monomial_tw, n, ilog2(n), ilog2(plan->base_n),
                                                    uint32_t(degree));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_permute(const cfft_plan *plan, bool to_standard, const double2 *src, double2 *dst, uint64_t batch,
                           cudaStream_t stream)
{
    const uint64_t total = plan->n * batch;
    if (total == 0) return cudaSuccess;
    uint64_t blocks = (total + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    const uint32_t nbits = ilog2(plan->n), bbits = plan->kind == KIND_ORDERED ? nbits : ilog2(plan->base_n);
    if (to_standard)
        permute_kernel<true><<<unsigned(blocks), 256, 0, stream>>>(src, dst, total, uint32_t(plan->n), nbits, bbits);
    else
        permute_kernel<false><<<unsigned(blocks), 256, 0, stream>>>(src, dst, total, uint32_t(plan->n), nbits, bbits);
    count_launch();
    return cudaGetLastError();
}

} // namespace cfft

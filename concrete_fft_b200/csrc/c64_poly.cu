// c64_poly.cu -- integer polynomials in, integer polynomials out: the caller-side steps around the c64 transform fused
// into its first / last pass.
//
// A negacyclic product of real polynomials mod X^N + 1 through an fft of size n = N / 2 is: fold the N coefficients into
// n complex points (re = coeff[j], im = coeff[j + n], src/fft128/mod.rs:2006-2016), convert to f64, twist by
// e^{+i pi j / N}, forward transform, element-wise products in the Fourier domain (README.md:10-17), inverse transform,
// untwist, scale by 1 / n, round back to integers.  Done with separate kernels the conversions cost two more trips through
// HBM than the transforms themselves; here they ride on the loads of the first butterfly level and on the stores of the
// last one (RowIo<PIN, POUT> in c64_dev.cuh), for every (Dif16, 256) plan with 256 <= n <= 8192 -- TFHE's polynomial sizes
// 512 .. 16384.  Other plans run the same arithmetic as stand-alone conversion kernels around their own transform.
// Semantics of the conversions (twist table from sincospi64, num_complex product, f64::round, torus = 2^-64 scaling):
// include/cfft_b200.h; the transforms in between are the reference's, bit for bit.
#include "c64_fast_kernels.cuh"

namespace cfft {
using namespace dev;
using namespace fastk;
namespace {

BatchIo<true, false> poly_in_batch(const long long *pin, uint64_t prow_in, c64 *out, uint64_t row_out, const c64 *twist, uint32_t n, uint32_t flags)
{
    BatchIo<true, false> b;
    b.in = nullptr;
    b.out = out;
    b.pin = pin;
    b.pout = nullptr;
    b.twist = twist;
    b.row_in = 0;
    b.row_out = row_out;
    b.prow_in = prow_in;
    b.prow_out = 0;
    b.n = n;
    b.flags = flags;
    b.ahead = 0;
    return b;
}
BatchIo<false, true> poly_out_batch(const c64 *in, uint64_t row_in, long long *pout, uint64_t prow_out, const c64 *twist, uint32_t n, uint32_t flags)
{
    BatchIo<false, true> b;
    b.in = in;
    b.out = nullptr;
    b.pin = nullptr;
    b.pout = pout;
    b.twist = twist;
    b.row_in = row_in;
    b.row_out = 0;
    b.prow_in = 0;
    b.prow_out = prow_out;
    b.n = n;
    b.flags = flags;
    b.ahead = 0;
    return b;
}

// stand-alone conversions (plans without a fused kernel): exactly RowIo's load / store, one element per thread
__global__ void __launch_bounds__(256) poly_fold_twist_kernel(BatchIo<true, false> bio, uint64_t rows)
{
    const uint64_t total = rows * bio.n;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t r = i / bio.n;
        const int pos = int(i - r * bio.n);
        const RowIo<true, false> io = bio.row(r);
        io.out[pos] = io.ld(pos);
    }
}
__global__ void __launch_bounds__(256) poly_untwist_round_kernel(BatchIo<false, true> bio, uint64_t rows)
{
    const uint64_t total = rows * bio.n;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
        const uint64_t r = i / bio.n;
        const int pos = int(i - r * bio.n);
        const RowIo<false, true> io = bio.row(r);
        io.st(pos, io.in[pos]);
    }
}
unsigned stream_grid(uint64_t elems)
{
    uint64_t blocks = (elems + 255) / 256;
    return unsigned(blocks > 148ull * 32 ? 148ull * 32 : (blocks ? blocks : 1));
}

template <int N, int R1, int R2>
cudaError_t launch_fwd_poly(const cfft_plan *plan, const long long *poly, c64 *fourier, uint64_t batch, uint32_t flags, cudaStream_t stream)
{
    using Cfg = FastCfg<N>;
    const size_t smem = size_t(Cfg::ROWS) * N * sizeof(c64);
    auto k = c64_fast_b256_kernel<N, R1, R2, true, false, true, false>;
    cudaError_t e = allow_smem(k, smem);
    if (e != cudaSuccess) return e;
    const uint64_t ctas = (batch + Cfg::ROWS - 1) / Cfg::ROWS;
    k<<<unsigned(ctas), Cfg::NT, smem, stream>>>(poly_in_batch(poly, 2 * N, fourier, N, plan->d_twist, N, flags), batch, fast_tables(plan, 0));
    count_launch();
    return cudaGetLastError();
}
template <int N, int R1, int R2>
cudaError_t launch_inv_poly(const cfft_plan *plan, const c64 *fourier, long long *poly, uint64_t batch, uint32_t flags, cudaStream_t stream)
{
    using Cfg = FastCfg<N>;
    const size_t smem = size_t(Cfg::ROWS) * N * sizeof(c64);
    auto k = c64_fast_b256_kernel<N, R1, R2, false, false, false, true>;
    cudaError_t e = allow_smem(k, smem);
    if (e != cudaSuccess) return e;
    const uint64_t ctas = (batch + Cfg::ROWS - 1) / Cfg::ROWS;
    k<<<unsigned(ctas), Cfg::NT, smem, stream>>>(poly_out_batch(fourier, N, poly, 2 * N, plan->d_twist, N, flags), batch, fast_tables(plan, 1));
    count_launch();
    return cudaGetLastError();
}
template <int N, int R1, int R2, bool ALLOW_MULTI>
cudaError_t launch_mul_poly(const cfft_plan *plan, const long long *a, uint32_t kterms, const c64 *b, uint64_t b_row_stride, long long *out,
                            uint64_t batch, uint32_t flags, cudaStream_t stream)
{
    using Cfg = FastCfg<N>;
    const size_t smem = size_t(Cfg::ROWS) * N * sizeof(c64);
    auto k1 = c64_fwd_mul_inv_kernel<N, R1, R2, false, true, true>;
    auto km = c64_fwd_mul_inv_kernel<N, R1, R2, ALLOW_MULTI, true, true>;
    cudaError_t e = allow_smem(k1, smem);
    if (e == cudaSuccess) e = allow_smem(km, smem);
    if (e != cudaSuccess) return e;
    const uint64_t ctas = (batch + Cfg::ROWS - 1) / Cfg::ROWS;
    const BatchIo<true, false> ain = poly_in_batch(a, 2 * N, nullptr, 0, plan->d_twist, N, flags);
    const BatchIo<false, true> oout = poly_out_batch(nullptr, 0, out, 2 * N, plan->d_twist, N, flags);
    // kernel flag bits 0 / 1: L2 prefetch of b / of the next term (as the plain fused product); the conversion flags travel in the accessors
    if (kterms == 1) k1<<<unsigned(ctas), Cfg::NT, smem, stream>>>(ain, b, oout, batch, kterms, b_row_stride, fast_tables(plan, 0), fast_tables(plan, 1), 1u);
    else km<<<unsigned(ctas), Cfg::NT, smem, stream>>>(ain, b, oout, batch, kterms, b_row_stride, fast_tables(plan, 0), fast_tables(plan, 1), 3u);
    count_launch();
    return cudaGetLastError();
}

bool has_fused_poly_kernels(const cfft_plan *plan)
{
    return plan->kind == KIND_UNORDERED && plan->d_fast_tw[0] && plan->n >= 256 && plan->n <= 8192 && plan->fast_variant != 0 && plan->fast_variant != 6 &&
           getenv("CFFT_B200_POLY_COMPOSED") == nullptr;
}

} // namespace

bool poly_fused_available(const cfft_plan *plan, uint64_t kterms)
{
    // n = 8192 (512 threads x 128 registers) has no room for a running sum: one term only
    return has_fused_poly_kernels(plan) && (plan->n <= 4096 || kterms <= 1);
}

// fourier[r] = fwd(twist(fold(poly[r]))): 2n coefficients in, n Fourier coefficients (this plan's order) out
cudaError_t launch_c64_poly_fwd(const cfft_plan *plan, const long long *poly, double2 *fourier, uint64_t batch, uint32_t flags, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    if (has_fused_poly_kernels(plan)) {
        switch (plan->n) {
        case 256: return launch_fwd_poly<256, 1, 1>(plan, poly, fourier, batch, flags, stream);
        case 512: return launch_fwd_poly<512, 2, 1>(plan, poly, fourier, batch, flags, stream);
        case 1024: return launch_fwd_poly<1024, 4, 1>(plan, poly, fourier, batch, flags, stream);
        case 2048: return launch_fwd_poly<2048, 8, 1>(plan, poly, fourier, batch, flags, stream);
        case 4096: return launch_fwd_poly<4096, 8, 2>(plan, poly, fourier, batch, flags, stream);
        case 8192: return launch_fwd_poly<8192, 8, 4>(plan, poly, fourier, batch, flags, stream);
        default: break;
        }
    }
    const uint32_t n = uint32_t(plan->n);
    poly_fold_twist_kernel<<<stream_grid(batch * n), 256, 0, stream>>>(poly_in_batch(poly, 2ull * n, fourier, n, plan->d_twist, n, flags), batch);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = launch_c64(plan, false, fourier, batch, stream);
    return e;
}

// poly[r] (+)= round(untwist(inv(fourier[r]))); fourier is left untouched
cudaError_t launch_c64_poly_inv(const cfft_plan *plan, const double2 *fourier, long long *poly, uint64_t batch, uint32_t flags, cudaStream_t stream)
{
    if (batch == 0) return cudaSuccess;
    if (has_fused_poly_kernels(plan)) {
        switch (plan->n) {
        case 256: return launch_inv_poly<256, 1, 1>(plan, fourier, poly, batch, flags, stream);
        case 512: return launch_inv_poly<512, 2, 1>(plan, fourier, poly, batch, flags, stream);
        case 1024: return launch_inv_poly<1024, 4, 1>(plan, fourier, poly, batch, flags, stream);
        case 2048: return launch_inv_poly<2048, 8, 1>(plan, fourier, poly, batch, flags, stream);
        case 4096: return launch_inv_poly<4096, 8, 2>(plan, fourier, poly, batch, flags, stream);
        case 8192: return launch_inv_poly<8192, 8, 4>(plan, fourier, poly, batch, flags, stream);
        default: break;
        }
    }
    // composed: the plan's inverse works in place, so it runs on a copy, slice by slice (<= 256 MiB of workspace)
    const uint32_t n = uint32_t(plan->n);
    const uint64_t row_bytes = uint64_t(n) * sizeof(c64);
    uint64_t chunk_rows = (uint64_t{256} << 20) / row_bytes;
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > batch) chunk_rows = batch;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(plan->device, &pool);
    if (e != cudaSuccess) return e;
    c64 *ws = nullptr;
    if ((e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), chunk_rows * row_bytes, pool, stream)) != cudaSuccess) return e;
    for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
        const uint64_t rows = batch - r0 < chunk_rows ? batch - r0 : chunk_rows;
        e = cudaMemcpyAsync(ws, fourier + r0 * n, rows * row_bytes, cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = launch_c64(plan, true, ws, rows, stream);
        if (e != cudaSuccess) break;
        poly_untwist_round_kernel<<<stream_grid(rows * n), 256, 0, stream>>>(poly_out_batch(ws, n, poly + r0 * 2 * n, 2ull * n, plan->d_twist, n, flags), rows);
        count_launch();
        e = cudaGetLastError();
    }
    const cudaError_t e2 = cudaFreeAsync(ws, stream);
    return e != cudaSuccess ? e : e2;
}

// out[r] (+)= round(untwist(inv(sum_k fwd(twist(fold(a[r][k]))) (.) b[r][k]))): a whole negacyclic product step, integers in
// and out.  a: [batch][kterms][2n] coefficients, b: Fourier-domain c64 ([kterms][n] shared when b_row_stride == 0), out: [batch][2n].
cudaError_t launch_c64_poly_mul(const cfft_plan *plan, const long long *a, uint64_t kterms, const double2 *b, uint64_t b_row_stride,
                                long long *out, uint64_t batch, uint32_t flags, cudaStream_t stream)
{
    if (batch == 0 || kterms == 0) return cudaSuccess;
    if (poly_fused_available(plan, kterms) && kterms <= 0xFFFFFFFFull) {
        const uint32_t kt = uint32_t(kterms);
        switch (plan->n) {
        case 256: return launch_mul_poly<256, 1, 1, true>(plan, a, kt, b, b_row_stride, out, batch, flags, stream);
        case 512: return launch_mul_poly<512, 2, 1, true>(plan, a, kt, b, b_row_stride, out, batch, flags, stream);
        case 1024: return launch_mul_poly<1024, 4, 1, true>(plan, a, kt, b, b_row_stride, out, batch, flags, stream);
        case 2048: return launch_mul_poly<2048, 8, 1, true>(plan, a, kt, b, b_row_stride, out, batch, flags, stream);
        case 4096: return launch_mul_poly<4096, 8, 2, true>(plan, a, kt, b, b_row_stride, out, batch, flags, stream);
        case 8192: return launch_mul_poly<8192, 8, 4, false>(plan, a, kt, b, b_row_stride, out, batch, flags, stream);
        default: break;
        }
    }
    // composed: fold + twist into a workspace, the plan's fused / composed product, untwist + round -- slice by slice
    const uint32_t n = uint32_t(plan->n);
    const uint64_t row_bytes = uint64_t(n) * sizeof(c64);
    uint64_t chunk_rows = (uint64_t{256} << 20) / (row_bytes * (kterms + 1));
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > batch) chunk_rows = batch;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(plan->device, &pool);
    if (e != cudaSuccess) return e;
    c64 *ws = nullptr;
    if ((e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), chunk_rows * (kterms + 1) * row_bytes, pool, stream)) != cudaSuccess) return e;
    c64 *ws_out = ws + chunk_rows * kterms * n;
    for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
        const uint64_t rows = batch - r0 < chunk_rows ? batch - r0 : chunk_rows;
        poly_fold_twist_kernel<<<stream_grid(rows * kterms * n), 256, 0, stream>>>(
            poly_in_batch(a + r0 * kterms * 2 * n, 2ull * n, ws, n, plan->d_twist, n, flags), rows * kterms);
        count_launch();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = launch_c64_fwd_mul_inv(plan, ws, kterms, b + r0 * b_row_stride, b_row_stride, ws_out, rows, stream);
        if (e != cudaSuccess) break;
        poly_untwist_round_kernel<<<stream_grid(rows * n), 256, 0, stream>>>(poly_out_batch(ws_out, n, out + r0 * 2 * n, 2ull * n, plan->d_twist, n, flags), rows);
        count_launch();
        e = cudaGetLastError();
    }
    const cudaError_t e2 = cudaFreeAsync(ws, stream);
    return e != cudaSuccess ? e : e2;
}

} // namespace cfft

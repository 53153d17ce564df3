// plan.h -- the plan object behind the C ABI (include/cfft_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "tables.h"

namespace cfft {

enum PlanKind { KIND_ORDERED = 0, KIND_UNORDERED = 1, KIND_F128 = 2 };

// One butterfly stage of the reference's schedule, in execution order.
enum StageKind : int {
    ST_TOP = 0,      // unordered in-place radix-2/4/8 level (src/unordered.rs:222-293); span = n_cur
    ST_CORE_DIF = 1, // Stockham DIF stage, twiddles on outputs (e.g. src/dif4.rs:118-168); span = s
    ST_CORE_DIT = 2, // Stockham DIT stage, twiddles on inputs  (e.g. src/dit4.rs:96-145);  span = s
    ST_END = 3       // terminal twiddle-free pass (e.g. src/dif4.rs:217-244)
};

struct Stage {
    int kind;
    int radix;
    uint32_t span;   // ST_TOP: n_cur; cores: stride s; END: unused
    uint32_t tw_off; // offset (in c64) of this stage's twiddles inside the direction's table
    uint32_t tw2;    // ST_TOP: offset of the planar copy w_k[p] (k-major) inside d_top_tw (c64_regs.cu)
};

constexpr int kMaxStages = 24;
struct StageProgram {
    int count;
    Stage st[kMaxStages];
};

constexpr uint32_t kTileMax = 4096; // c64 per CTA tile in the exact kernel (2 x 64 KiB of smem)

} // namespace cfft

// Opaque to C callers.
struct cfft_plan {
    int kind = 0;
    int device = 0;
    uint64_t n = 0;
    int algo = 0;        // ordered: algo; unordered: base_algo
    uint64_t base_n = 0; // ordered: n
    int method = 0;
    bool allow_large = false;
    std::string kernel_name;
    std::string tuning_report;
    uint32_t tile_elems = 0; // exact / fft128 tile size override chosen by the autotuner (0 = default)
    int f128_smax = 3;       // fft128 tile kernel: 3 = three-stage groups, 2 CTAs per SM; 2 = two-stage groups, 3 CTAs per SM
    // multi-pass (variant 2) scheduling: 0 = whole batch per pass; else passes run chunk by chunk
    // (chunk <= l2_chunk_mb MiB, alternating over l2_streams auxiliary streams) so that a chunk stays
    // L2-resident between its passes
    uint32_t l2_chunk_mb = 0, l2_streams = 1;

    // c64
    std::vector<cfft::cplx> h_tw[2]; // [0] fwd, [1] inv (host copies, kept for clone / tests)
    double2 *d_tw[2] = {nullptr, nullptr};
    double2 *d_monomial_tw = nullptr; // n entries, e^{-2 pi i k / n} (src/unordered.rs:714-720)
    double2 *d_twist = nullptr;       // 2n entries: twist e^{+i pi j / 2n}, then untwist conj(twist) / n (cfft_c64_poly_*)
    cfft::StageProgram prog[2];       // [0] fwd, [1] inv: every stage in execution order
    double2 *d_top_tw[2] = {nullptr, nullptr}; // planar copies of the unordered level tables (same values)
    bool exact_regs = true;           // plans without a specialised kernel: register kernel (c64_regs.cu) or tile kernel
    int fast_variant = 0;             // 0 exact tile kernel; 1 fused register kernel; 2 column passes + rows;
                                      // 3 ordered (standard order in/out) above 2^10: column passes + transposing rows
                                      // 4 one transform per thread-block cluster (n = 8192, 16384), DSMEM exchange
                                      // 5 ordered above 2^10, n <= 8192: fused register kernel with standard-order in / out
                                      // 9 n >= 2^14: column passes for the upper levels + fused kernel of 512 .. 4096 points for the rest
                                      // 8 n = 2^14 .. 2^16: both HBM passes in one persistent kernel, intermediate kept in L2
                                      // 6 whole-transform Dif16 plans, n = 32..128, 512, 1024 (c64_ord16.cu)
    double2 *d_fast_tw[2] = {nullptr, nullptr}; // planar re-layout of the same twiddle values
    struct FastLevel { int radix; uint32_t span; uint32_t off; }; // off: planar table inside d_fast_tw
    std::vector<FastLevel> fast_levels;         // unordered levels, outermost first
    uint32_t fast_base_off = 0;                 // planar half of the base init_wt table
    struct FastGroup { int radices[3]; uint32_t span0; int first_level; };
    std::vector<FastGroup> fast_groups;         // fast_variant == 2: levels grouped per HBM pass
    // fast_variant == 9: the last one or two levels + the base FFTs run as the fused kernel of size tail_n
    // (512 .. 4096) on contiguous blocks; only the levels above it are column passes (one HBM pass fewer for n >= 2^17)
    std::vector<FastGroup> tail_groups;
    uint32_t tail_n = 0;
    int tail_first_level = 0;

    // fft128
    std::vector<double> h_f128_tw[4];
    double *d_f128_tw[4] = {nullptr, nullptr, nullptr, nullptr};
    double *d_f128_tw4 = nullptr; // same values interleaved {re hi, re lo, im hi, im lo} per index
};

namespace cfft {

// kernels (c64_tile.cu)
cudaError_t launch_c64_exact(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t st);
cudaError_t launch_monomial(const cfft_plan *plan, uint64_t degree, double2 *data, cudaStream_t st);
cudaError_t launch_c64_pointwise(double2 *acc, double2 *a, const double2 *b, uint64_t len, cudaStream_t st);
cudaError_t launch_permute(const cfft_plan *plan, bool to_standard, const double2 *src, double2 *dst,
                           uint64_t batch, cudaStream_t st);
// kernels (c64_regs.cu): any stage program on a tile of 2048 / 4096 elements
cudaError_t launch_c64_regs(bool inverse, uint32_t tile, const StageProgram &prog, double2 *data, uint64_t total,
                            uint32_t base_n, const double2 *tw_ref, const double2 *tw_top, cudaStream_t st);
// the same kernel with the schedule of a few common unordered plans built at compile time (no index arithmetic left)
bool regs_spec_supported(uint64_t n, int radix, bool dit, uint64_t base_n);
cudaError_t launch_c64_regs_spec(bool inverse, uint64_t n, int radix, bool dit, uint64_t base_n, const StageProgram &prog, double2 *data,
                                 uint64_t total, const double2 *tw_ref, const double2 *tw_top, cudaStream_t st, bool *taken);
// kernels (c64_fast.cu)
bool fast_b256_supported(uint64_t n, int base_algo, uint64_t base_n);
cudaError_t launch_c64_fast_b256(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t st);
bool fast_b256_strided_available(const cfft_plan *plan);
cudaError_t launch_c64_fast_b256_strided(const cfft_plan *plan, bool inverse, double2 *data, uint64_t row_stride, uint64_t batch,
                                         cudaStream_t st);
// fwd -> point-wise multiply-accumulate over k terms -> inv (c64_fast.cu): one fused kernel for (Dif16, 256) plans
// with n <= 4096, the same arithmetic composed from the plan's kernels otherwise
bool fused_mul_kernel_available(const cfft_plan *plan);
cudaError_t launch_c64_fwd_mul_inv(const cfft_plan *plan, const double2 *a, uint64_t kterms, const double2 *b,
                                   uint64_t b_row_stride, double2 *out, uint64_t batch, cudaStream_t st);
// the same with n_out outputs per row sharing the forward transforms (b: [k][n_out][n]); one kernel for n_out == 2, n = 512 .. 2048
bool fused_mul2_kernel_available(const cfft_plan *plan);
cudaError_t launch_c64_fwd_mul_inv_multi(const cfft_plan *plan, const double2 *a, uint64_t kterms, const double2 *b, uint64_t b_row_stride,
                                         uint64_t n_out, double2 *out, uint64_t batch, cudaStream_t st);
cudaError_t launch_c64_fwd_mul_add(const cfft_plan *plan, const double2 *a, uint64_t a_row_terms, const double2 *b,
                                   uint64_t b_row_stride, double2 *acc, bool accumulate, uint64_t batch, cudaStream_t st);
// kernels (c64_poly.cu): integer polynomials (2n signed 64-bit coefficients per row) <-> the Fourier domain with the fold,
// the conversion, the negacyclic twist and the rounding fused into the transform's first / last pass; flags: bit 0 torus
// (x 2^-64 in, fractional part x 2^64 out), bit 1 accumulate into the output polynomial (modulo 2^64)
bool poly_fused_available(const cfft_plan *plan, uint64_t kterms);
cudaError_t launch_c64_poly_fwd(const cfft_plan *plan, const long long *poly, double2 *fourier, uint64_t batch, uint32_t flags, cudaStream_t st);
cudaError_t launch_c64_poly_inv(const cfft_plan *plan, const double2 *fourier, long long *poly, uint64_t batch, uint32_t flags, cudaStream_t st);
cudaError_t launch_c64_poly_mul(const cfft_plan *plan, const long long *a, uint64_t kterms, const double2 *b, uint64_t b_row_stride,
                                long long *out, uint64_t batch, uint32_t flags, cudaStream_t st);
// kernels (c64_ord16.cu)
bool ord16_supported(uint64_t n, int algo);
cudaError_t launch_c64_ord16(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t st);
// kernels (c64_column.cu): one group of <= 3 unordered levels in one HBM pass
cudaError_t launch_c64_column_group(bool inverse, const double2 *src, double2 *dst, uint64_t batch, uint32_t n, uint32_t span0,
                                    const int radices[3], const double2 *const tw[3], cudaStream_t st);
// kernels (c64_colpipe.cu): the same groups as a persistent kernel -- a CTA keeps one tile position and walks through the
// batch with the position's twiddles in registers and the next tile's loads in flight
bool colpipe_supported(const int radices[3]);
cudaError_t launch_c64_colpipe_group(bool inverse, const double2 *src, double2 *dst, uint64_t batch, uint32_t n, uint32_t span0,
                                     const int radices[3], const double2 *const tw[3], int device, cudaStream_t st);
// kernels (c64_tmem.cu): two radix-8 levels in one HBM pass, one thread per 64-element column, tensor memory as the
// parking space between the levels (no shared memory, no block barrier)
bool tmem_column88_supported(uint32_t n, uint32_t span0);
cudaError_t launch_c64_tmem_column88(bool inverse, const double2 *src, double2 *dst, uint64_t batch, uint32_t n, uint32_t span0,
                                     const double2 *tw0, const double2 *tw1, cudaStream_t st);
// n = 2^14 .. 2^16: column group + base FFTs in one persistent kernel (c64_column.cu)
cudaError_t launch_c64_twopass(bool inverse, double2 *data, uint64_t batch, uint32_t n, const int radices[3],
                               const double2 *const tw[3], const double2 *tw_base, uint32_t lag, int device, cudaStream_t st);
cudaError_t twopass_timeouts(unsigned int *out);
// stream-ordered scratch pool, one per device, keeps its memory between calls (c64_fast.cu)
cudaError_t workspace_pool(int device, cudaMemPool_t *out);
// dispatcher (api.cc): fast kernel when the plan has one, else the exact tile kernel
cudaError_t launch_c64(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t st);
// kernels (f128.cu)
cudaError_t launch_f128(const cfft_plan *plan, bool inverse, double *re0, double *re1, double *im0, double *im1,
                        uint64_t batch, cudaStream_t st);

// kernels (f128_ops.cu)
cudaError_t launch_f128_binary(int op, const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo,
                               double *out_hi, double *out_lo, uint64_t len, cudaStream_t st);
cudaError_t launch_f128_unary(int op, const double *a_hi, const double *a_lo, double *out_hi, double *out_lo, double *out2_hi,
                              double *out2_lo, uint64_t len, unsigned int *bad, cudaStream_t st);
cudaError_t launch_f128_compare(const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo, signed char *out,
                                uint64_t len, cudaStream_t st);
cudaError_t launch_f128_cplx_mul_scale(double *l_re0, double *l_re1, double *l_im0, double *l_im1, const double *r_re0,
                                       const double *r_re1, const double *r_im0, const double *r_im1, double factor,
                                       uint64_t len, cudaStream_t st);

cudaError_t launch_f128_cplx_mul_scale_rows(double *l_re0, double *l_re1, double *l_im0, double *l_im1, const double *r_re0,
                                            const double *r_re1, const double *r_im0, const double *r_im1, uint64_t rhs_period,
                                            double factor, uint64_t len, cudaStream_t st);
// fft128 fwd -> point-wise product * factor -> inv, one kernel for n <= 4096 (f128.cu)
bool f128_fused_mul_kernel_available(const cfft_plan *plan);
cudaError_t launch_f128_fwd_mul_inv(const cfft_plan *plan, double *l_re0, double *l_re1, double *l_im0, double *l_im1,
                                    const double *r_re0, const double *r_re1, const double *r_im0, const double *r_im1,
                                    bool rhs_shared, double factor, uint64_t batch, cudaStream_t st);

void count_launch(uint64_t k = 1);

} // namespace cfft

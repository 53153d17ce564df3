// c64_regs.cu -- register-resident execution of ANY plan's stage schedule (all 8 algorithms, every
// base_n, ordered and unordered): the successor of the shared-memory "exact tile" kernel of c64_tile.cu.
//
// A CTA of NT threads owns a tile of 16 NT contiguous c64 (whole transforms, or a sub-block of one when
// n exceeds the tile) and every thread always holds 16 of them in registers: a radix-R stage is 16/R
// butterflies per thread, so EVERY stage of the reference's schedule keeps all threads busy --
//   unordered levels   fwd_process_x{2,4,8} / inv_process_x{2,4,8}     src/unordered.rs:222-293
//   Stockham stages    stockham_core_* of dif{2,4,8,16}.rs / dit{2,4,8,16}.rs (e.g. src/dif4.rs:118-168,
//                      src/dit4.rs:96-145) and the terminal passes (e.g. src/dif4.rs:217-244)
// -- with the first stage reading HBM and the last writing HBM directly (their access patterns are
// lane-consecutive for every stage kind), stages in between exchanging through ONE swizzled shared-memory
// copy of the tile (read all -> barrier -> compute -> write, so no ping-pong buffer: 32 KiB per 128
// threads, four CTAs per SM), and twiddles read from planar tables (lanes on consecutive p) instead of
// the reference's per-butterfly interleaved layout.  Same butterflies (c64_math.cuh), same table values,
// same element order => bit-identical to the reference for the plan, like the kernel it replaces.
#include "c64_dev.cuh"
#include "plan.h"

namespace cfft {
using namespace dev;
namespace {

// conflict-free for every stage pattern (consecutive elements, or elements 2..16 apart across lanes)
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u) ^ ((i >> 6) & 7u); }

template <int R> __device__ __forceinline__ int brev_r(int k)
{
    return R == 16 ? ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | (k >> 3) : brev_c<R>(k);
}

struct TileIo {
    c64 *g;         // tile start in HBM
    c64 *s;         // tile in shared memory (swizzled)
    uint32_t valid; // elements of the tile that exist (whole transforms)
    bool in_g, out_g;
};

// One stage on the 16 register values of a thread.  in_idx / out_idx: tile positions of element k of
// butterfly j; tw_idx: twiddle index of (j, k >= 1) or nullptr-equivalent when the stage has none.
// A partial tile (the last CTA of a ragged batch) is handled without per-butterfly liveness: loads from
// HBM clamp the index into the valid range, butterflies past the end compute on garbage that only ever lands
// in shared memory, and only the HBM stores of the last stage are guarded.
template <int R, bool FWD, bool TW_IN, bool TW_OUT, class FIn, class FOut, class FTw>
__device__ __forceinline__ void stage_body(const TileIo &io, const c64 *__restrict__ tw, FIn in_idx, FOut out_idx, FTw tw_idx,
                                           c64 (&v)[16])
{
    constexpr int B = 16 / R;
    if (io.in_g) {
        const uint32_t last = io.valid - 1;
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) v[j * R + k] = ld_stream(io.g + min(in_idx(j, k), last));
    } else {
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) v[j * R + k] = io.s[swz(in_idx(j, k))];
        if (!io.out_g) __syncthreads(); // in place: every read before any write
    }
#pragma unroll
    for (int j = 0; j < B; j++) {
        c64 *x = &v[j * R];
        if (TW_IN) {
#pragma unroll
            for (int k = 1; k < R; k++) x[k] = cmul(ld_tw(tw + tw_idx(j, k)), x[k]);
        }
        bfR<R, FWD>(x);
        if (TW_OUT) {
#pragma unroll
            for (int k = 1; k < R; k++) x[k] = cmul(ld_tw(tw + tw_idx(j, k)), x[k]);
        }
    }
    if (io.out_g) {
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) {
                const uint32_t i = out_idx(j, k);
                if (i < io.valid) st_stream(io.g + i, v[j * R + k]);
            }
    } else {
#pragma unroll
        for (int j = 0; j < B; j++)
#pragma unroll
            for (int k = 0; k < R; k++) io.s[swz(out_idx(j, k))] = v[j * R + k];
    }
}

__device__ __forceinline__ uint32_t lg2(uint32_t x) { return 31u - uint32_t(__clz(int(x))); }

// Index maps are recomputed from the butterfly number b = t + NT j where they are used (a few integer
// operations next to ~50 FP64 instructions per element) instead of being kept in registers.
template <int R, bool FWD, int NT>
__device__ __forceinline__ void run_stage(const Stage &st, uint32_t base_n, const c64 *tw_ref, const c64 *tw_top,
                                          const TileIo &io, c64 (&v)[16])
{
    const uint32_t t = threadIdx.x;
    if (st.kind == ST_TOP) {
        // fwd: v = DFT_R(z[p + m k]); z[p + m brev(k)] = w_k v_k      inv: v_k = w_k z[p + m brev(k)]; z[p + m k] = DFT^-1
        const uint32_t m = st.span / R, lm = lg2(m), span = st.span;
        auto zof = [=](int j) { const uint32_t b = t + NT * j; return (b >> lm) * span + (b & (m - 1)); };
        auto nat = [=](int j, int k) { return zof(j) + m * uint32_t(k); };
        auto rev = [=](int j, int k) { return zof(j) + m * uint32_t(brev_r<R>(k)); };
        auto twi = [=](int j, int k) { return uint32_t(k - 1) * m + ((t + NT * j) & (m - 1)); };
        if (FWD) stage_body<R, true, false, true>(io, tw_top + st.tw2, nat, rev, twi, v);
        else stage_body<R, false, true, false>(io, tw_top + st.tw2, rev, nat, twi, v);
    } else if (st.kind == ST_END) {
        const uint32_t part = base_n / R, lp = lg2(part);
        auto idx = [=](int j, int k) { const uint32_t b = t + NT * j; return (b >> lp) * base_n + (b & (part - 1)) + part * uint32_t(k); };
        auto twi = [](int, int) { return 0u; };
        stage_body<R, FWD, false, false>(io, tw_ref, idx, idx, twi, v);
    } else {
        // Stockham stage at stride s: natural side x[q + s(p + m k)], scattered side y[q + s(R p + k)], twiddle w^{k p s}
        const uint32_t s = st.span, per_blk = base_n / R, lpb = lg2(per_blk), ls = lg2(s);
        auto nat = [=](int j, int k) {
            const uint32_t b = t + NT * j;
            return (b >> lpb) * base_n + (b & (per_blk - 1)) + per_blk * uint32_t(k);
        };
        auto sca = [=](int j, int k) {
            const uint32_t b = t + NT * j, rem = b & (per_blk - 1);
            return (b >> lpb) * base_n + (rem & (s - 1)) + (((rem >> ls) * R + uint32_t(k)) << ls);
        };
        auto twi = [=](int j, int k) { // planar half: w[p s + k base_n / R]
            const uint32_t rem = (t + NT * j) & (per_blk - 1);
            return (rem & ~(s - 1)) + per_blk * uint32_t(k);
        };
        const c64 *w = tw_ref + st.tw_off - base_n;
        if (st.kind == ST_CORE_DIF) stage_body<R, FWD, false, true>(io, w, nat, sca, twi, v);
        else stage_body<R, FWD, true, false>(io, w, sca, nat, twi, v);
    }
}

// (Measured and rejected: a ping-pong variant -- two copies of the tile, one barrier per stage, three CTAs per
// SM -- 10-25 % slower; 168 registers / 12 warps per SM to avoid the ~270 B of spills -- no better on balance.)
template <bool FWD, int NT, int WPS = 16>
__global__ void __launch_bounds__(NT, WPS * 32 / NT)
c64_regs_kernel(c64 *__restrict__ data, uint64_t total, uint32_t base_n, StageProgram prog, const c64 *__restrict__ tw_ref,
                const c64 *__restrict__ tw_top)
{
    constexpr uint32_t TILE = 16 * NT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileIo io;
    io.s = reinterpret_cast<c64 *>(smem_raw);
    const uint64_t start = uint64_t(blockIdx.x) * TILE;
    io.g = data + start;
    io.valid = total - start < TILE ? uint32_t(total - start) : TILE;
    c64 v[16];
    for (int si = 0; si < prog.count; si++) {
        const Stage st = prog.st[si];
        io.in_g = si == 0;
        io.out_g = si == prog.count - 1;
        if (!io.in_g) __syncthreads(); // the previous stage's shared-memory writes
        if (st.radix == 16) run_stage<16, FWD, NT>(st, base_n, tw_ref, tw_top, io, v);
        else if (st.radix == 8) run_stage<8, FWD, NT>(st, base_n, tw_ref, tw_top, io, v);
        else if (st.radix == 4) run_stage<4, FWD, NT>(st, base_n, tw_ref, tw_top, io, v);
        else run_stage<2, FWD, NT>(st, base_n, tw_ref, tw_top, io, v);
    }
}

// ---- compile-time stage schedules --------------------------------------------------------------------------------
// The interpreter above spends ~45 % of its instructions on index arithmetic for stage shapes it only learns at run time
// (profiles/r1z_ncu_regs_*).  For the plans the reference's own Method::Measure tends to produce (base_n = 512 / 1024 with
// the radix-8 / 16 algorithms, src/unordered.rs:568-630) and the golden-vector plan (Dif4, 32) the schedule is built at
// COMPILE time by the same rules as build_c64_programs / build_top_planar (api.cc), the stage loop is unrolled and every
// stride, mask and table offset folds into an immediate.  The launcher compares the compile-time schedule with the
// plan's own stage by stage and only then takes this path, so the two can never diverge.  Same run_stage, same bits.
struct CProg {
    int count;
    Stage st[12];
};

constexpr uint32_t c_ilog2(uint32_t x) { return x <= 1 ? 0 : 1 + c_ilog2(x >> 1); }

constexpr void c_append_base(CProg &pg, int R, bool dit, uint32_t base_n, uint32_t tw_off)
{
    const uint32_t rho = c_ilog2(uint32_t(R));
    uint32_t bits = c_ilog2(base_n), s = 1, strides[12] = {}, ns = 0;
    while (bits > rho) {
        strides[ns++] = s;
        s *= uint32_t(R);
        bits -= rho;
    }
    if (!dit) {
        for (uint32_t i = 0; i < ns; i++) pg.st[pg.count++] = Stage{ST_CORE_DIF, R, strides[i], tw_off, 0};
        pg.st[pg.count++] = Stage{ST_END, 1 << bits, 0, tw_off, 0};
    } else {
        pg.st[pg.count++] = Stage{ST_END, 1 << bits, 0, tw_off, 0};
        for (uint32_t i = ns; i-- > 0;) pg.st[pg.count++] = Stage{ST_CORE_DIT, R, strides[i], tw_off, 0};
    }
}

constexpr CProg make_cprog(uint32_t n, int R, bool dit, uint32_t base_n, bool inverse)
{
    CProg pg = {};
    struct Lvl { int r; uint32_t span, off_f, off_i; } lv[8] = {};
    int nl = 0;
    uint32_t cur = n, head = 0, tail = n + base_n;
    while (cur > base_n) {
        const int r = cur == 2 * base_n ? 2 : (cur == 4 * base_n ? 4 : 8);
        const uint32_t sz = uint32_t(r - 1) * (cur / uint32_t(r));
        tail -= sz;
        lv[nl++] = Lvl{r, cur, head, tail};
        head += sz;
        cur /= uint32_t(r);
    }
    uint32_t planar = 0;
    if (!inverse) {
        for (int i = 0; i < nl; i++) {
            pg.st[pg.count++] = Stage{ST_TOP, lv[i].r, lv[i].span, lv[i].off_f, planar};
            planar += uint32_t(lv[i].r - 1) * (lv[i].span / uint32_t(lv[i].r));
        }
        c_append_base(pg, R, dit, base_n, head + base_n);
    } else {
        c_append_base(pg, R, dit, base_n, base_n);
        for (int i = nl; i-- > 0;) {
            pg.st[pg.count++] = Stage{ST_TOP, lv[i].r, lv[i].span, lv[i].off_i, planar};
            planar += uint32_t(lv[i].r - 1) * (lv[i].span / uint32_t(lv[i].r));
        }
    }
    return pg;
}

template <bool FWD, uint32_t N, int R, bool DIT, uint32_t BASE_N> struct SpecProg {
    static constexpr CProg P = make_cprog(N, R, DIT, BASE_N, !FWD);
};

template <class SP, int SI, bool FWD, int NT>
__device__ __forceinline__ void run_spec_from(uint32_t base_n, const c64 *tw_ref, const c64 *tw_top, TileIo &io, c64 (&v)[16])
{
    constexpr Stage st = SP::P.st[SI];
    io.in_g = SI == 0;
    io.out_g = SI == SP::P.count - 1;
    if (SI != 0) __syncthreads(); // the previous stage's shared-memory writes
    run_stage<st.radix, FWD, NT>(st, base_n, tw_ref, tw_top, io, v);
    if constexpr (SI + 1 < SP::P.count) run_spec_from<SP, SI + 1, FWD, NT>(base_n, tw_ref, tw_top, io, v);
}

template <bool FWD, uint32_t N, int R, bool DIT, uint32_t BASE_N>
__global__ void __launch_bounds__((N < 1024 ? 1024 : N) / 16, 16 * 32 / ((N < 1024 ? 1024 : N) / 16))
c64_regs_spec_kernel(c64 *__restrict__ data, uint64_t total, const c64 *__restrict__ tw_ref, const c64 *__restrict__ tw_top)
{
    constexpr int NT = (N < 1024 ? 1024 : N) / 16;
    constexpr uint32_t TILE = 16 * NT;
    using SP = SpecProg<FWD, N, R, DIT, BASE_N>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileIo io;
    io.s = reinterpret_cast<c64 *>(smem_raw);
    const uint64_t start = uint64_t(blockIdx.x) * TILE;
    io.g = data + start;
    io.valid = total - start < TILE ? uint32_t(total - start) : TILE;
    c64 v[16];
    run_spec_from<SP, 0, FWD, NT>(BASE_N, tw_ref, tw_top, io, v);
}

inline bool same_program(const CProg &c, const StageProgram &p)
{
    if (c.count != p.count) return false;
    for (int i = 0; i < c.count; i++) {
        const Stage &a = c.st[i], &b = p.st[i];
        if (a.kind != b.kind || a.radix != b.radix || a.tw_off != b.tw_off) return false;
        if (a.kind != ST_END && a.span != b.span) return false;
        if (a.kind == ST_TOP && a.tw2 != b.tw2) return false;
    }
    return true;
}

template <bool FWD, uint32_t N, int R, bool DIT, uint32_t BASE_N>
cudaError_t launch_spec_t(const StageProgram &prog, c64 *data, uint64_t total, const c64 *tw_ref, const c64 *tw_top, cudaStream_t stream,
                          bool *taken)
{
    using SP = SpecProg<FWD, N, R, DIT, BASE_N>;
    constexpr int NT = (N < 1024 ? 1024 : N) / 16;
    constexpr uint32_t TILE = 16 * NT;
    static const CProg cp = SP::P;
    if (!same_program(cp, prog)) return cudaSuccess; // not this schedule after all: the interpreter runs it
    *taken = true;
    const size_t smem = size_t(TILE) * sizeof(c64);
    auto k = c64_regs_spec_kernel<FWD, N, R, DIT, BASE_N>;
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_device != dev) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            configured_device = dev;
        }
    }
    const uint64_t tiles = (total + TILE - 1) / TILE;
    k<<<unsigned(tiles), NT, smem, stream>>>(data, total, tw_ref, tw_top);
    count_launch();
    return cudaGetLastError();
}

template <bool FWD, int NT, int WPS = 16>
cudaError_t launch_regs_t(const StageProgram &prog, c64 *data, uint64_t total, uint32_t base_n, const c64 *tw_ref,
                          const c64 *tw_top, cudaStream_t stream)
{
    constexpr uint32_t TILE = 16 * NT;
    const size_t smem = size_t(TILE) * sizeof(c64);
    if (smem > 48 * 1024) {
        static thread_local int configured_device = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (configured_device != dev) {
            cudaError_t e = cudaFuncSetAttribute(c64_regs_kernel<FWD, NT, WPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            configured_device = dev;
        }
    }
    const uint64_t tiles = (total + TILE - 1) / TILE;
    c64_regs_kernel<FWD, NT, WPS><<<unsigned(tiles), NT, smem, stream>>>(data, total, base_n, prog, tw_ref, tw_top);
    count_launch();
    return cudaGetLastError();
}

} // namespace

// Plans with a compile-time schedule (whole transforms per tile; an ordered plan is the case base_n == n).  *taken tells
// whether the call was launched here.
#define CFFT_SPEC_PLANS(X)                                                                                          \
    X(2048, 16, false, 1024) X(2048, 16, false, 512) X(2048, 8, false, 512) X(2048, 4, false, 32) X(2048, 16, true, 1024) \
    X(1024, 16, false, 512) X(1024, 8, false, 512) X(4096, 16, false, 1024) X(4096, 8, false, 512)                  \
    X(2048, 8, true, 512) X(2048, 16, true, 512) X(1024, 16, true, 512) X(1024, 8, true, 512) X(4096, 16, true, 1024) X(4096, 8, true, 512) \
    /* whole-transform plans (ordered plans, and unordered plans with base_n == n): no levels, Stockham stages only */ \
    X(1024, 8, false, 1024) X(1024, 8, true, 1024) X(1024, 16, true, 1024) X(512, 8, false, 512) X(512, 8, true, 512) X(1024, 4, false, 1024)

bool regs_spec_supported(uint64_t n, int radix, bool dit, uint64_t base_n)
{
#define X(N, R, DIT, B) if (n == N && radix == R && dit == DIT && base_n == B) return true;
    CFFT_SPEC_PLANS(X)
#undef X
    return false;
}

cudaError_t launch_c64_regs_spec(bool inverse, uint64_t n, int radix, bool dit, uint64_t base_n, const StageProgram &prog, double2 *data,
                                 uint64_t total, const double2 *tw_ref, const double2 *tw_top, cudaStream_t stream, bool *taken)
{
    *taken = false;
    if (total == 0) return cudaSuccess;
#define X(N, R, DIT, B)                                                                                                          \
    if (n == N && radix == R && dit == DIT && base_n == B)                                                                       \
        return inverse ? launch_spec_t<false, N, R, DIT, B>(prog, data, total, tw_ref, tw_top, stream, taken)                     \
                       : launch_spec_t<true, N, R, DIT, B>(prog, data, total, tw_ref, tw_top, stream, taken);
    CFFT_SPEC_PLANS(X)
#undef X
    return cudaSuccess;
}

// tile: 1024 (64 threads), 2048 (128 threads) or 4096 (256 threads) elements; prog: the stages whose span fits the tile
cudaError_t launch_c64_regs(bool inverse, uint32_t tile, const StageProgram &prog, double2 *data, uint64_t total,
                            uint32_t base_n, const double2 *tw_ref, const double2 *tw_top, cudaStream_t stream)
{
    if (prog.count == 0 || total == 0) return cudaSuccess;
    if (tile == 4096)
        return inverse ? launch_regs_t<false, 256>(prog, data, total, base_n, tw_ref, tw_top, stream)
                       : launch_regs_t<true, 256>(prog, data, total, base_n, tw_ref, tw_top, stream);
    if (tile == 1024)
        return inverse ? launch_regs_t<false, 64>(prog, data, total, base_n, tw_ref, tw_top, stream)
                       : launch_regs_t<true, 64>(prog, data, total, base_n, tw_ref, tw_top, stream);
    return inverse ? launch_regs_t<false, 128>(prog, data, total, base_n, tw_ref, tw_top, stream)
                   : launch_regs_t<true, 128>(prog, data, total, base_n, tw_ref, tw_top, stream);
}

} // namespace cfft

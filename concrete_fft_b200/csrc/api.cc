// api.cc -- the extern "C" surface declared in include/cfft_b200.h: plan construction
// (tables + stage schedule), argument checking with the reference's preconditions, and the
// host-buffer pipelines.  No CPU fallback: every compute entry needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/cfft_b200.h"
#include "plan.h"

using namespace cfft;

namespace {

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};

cfft_status fail(cfft_status st, const std::string &msg)
{
    g_last_error = msg;
    return st;
}
cfft_status cuda_fail(cudaError_t e, const char *what)
{
    return fail(CFFT_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);           \
    } while (0)

// do the byte ranges [a, a + na) and [b, b + nb) intersect?
bool ranges_overlap(const void *a, uint64_t na, const void *b, uint64_t nb)
{
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(a), b0 = reinterpret_cast<uintptr_t>(b);
    return na && nb && a0 < b0 + nb && b0 < a0 + na;
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// ---- stage schedules ------------------------------------------------------------------------

// Stockham stages of the ordered base FFT of size base_n (type-level recursion of e.g.
// src/dif4.rs:246-303 / src/dit4.rs:223-280 flattened): appended in execution order.
void append_base_stages(StageProgram &pg, int algo, uint64_t base_n, uint32_t tw_off)
{
    if (base_n <= 1) return;
    const int R = algo_radix(algo);
    const unsigned rho = ilog2(uint64_t(R));
    unsigned bits = ilog2(base_n);
    std::vector<uint32_t> strides;
    uint32_t s = 1;
    while (bits > rho) {
        strides.push_back(s);
        s *= uint32_t(R);
        bits -= rho;
    }
    const Stage end = {ST_END, 1 << bits, 0, tw_off};
    if (!algo_is_dit(algo)) {
        for (uint32_t st : strides) pg.st[pg.count++] = Stage{ST_CORE_DIF, R, st, tw_off};
        pg.st[pg.count++] = end;
    } else {
        pg.st[pg.count++] = end;
        for (auto it = strides.rbegin(); it != strides.rend(); ++it) pg.st[pg.count++] = Stage{ST_CORE_DIT, R, *it, tw_off};
    }
}

void build_c64_programs(cfft_plan *p)
{
    const uint64_t n = p->n, base_n = p->base_n;
    StageProgram &f = p->prog[0], &v = p->prog[1];
    f.count = v.count = 0;

    // forward: unordered levels top-down (src/unordered.rs:392-439), then the base FFT
    struct Lvl { int r; uint64_t span; uint32_t off_f, off_i; };
    std::vector<Lvl> lv;
    uint64_t cur = n;
    uint32_t head = 0, tail = uint32_t(n + base_n);
    while (cur > base_n) {
        const int r = top_radix(cur, base_n);
        const uint32_t sz = uint32_t((r - 1) * (cur / r));
        tail -= sz;
        lv.push_back(Lvl{r, cur, head, tail});
        head += sz;
        cur /= r;
    }
    for (const Lvl &l : lv) f.st[f.count++] = Stage{ST_TOP, l.r, uint32_t(l.span), l.off_f};
    append_base_stages(f, p->algo, base_n, head + uint32_t(base_n));

    // inverse: base FFT first, then the levels bottom-up (src/unordered.rs:442-489)
    append_base_stages(v, p->algo, base_n, uint32_t(base_n));
    for (auto it = lv.rbegin(); it != lv.rend(); ++it) v.st[v.count++] = Stage{ST_TOP, it->r, uint32_t(it->span), it->off_i};
}

cfft_status upload_c64(cfft_plan *p)
{
    for (int d = 0; d < 2; d++) {
        const size_t bytes = p->h_tw[d].size() * sizeof(cplx);
        if (bytes == 0) continue;
        CU(cudaMalloc(reinterpret_cast<void **>(&p->d_tw[d]), bytes));
        CU(cudaMemcpy(p->d_tw[d], p->h_tw[d].data(), bytes, cudaMemcpyHostToDevice));
    }
    if (p->kind == KIND_UNORDERED) {
        // src/unordered.rs:714-720
        std::vector<cplx> mono(p->n);
        const double theta = -2.0 / double(p->n);
        for (uint64_t i = 0; i < p->n; i++) {
            double s, c;
            sincospi64(theta * double(i), s, c);
            mono[i] = cplx{c, s};
        }
        CU(cudaMalloc(reinterpret_cast<void **>(&p->d_monomial_tw), p->n * sizeof(cplx)));
        CU(cudaMemcpy(p->d_monomial_tw, mono.data(), p->n * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    {
        // negacyclic twist of the polynomial entry points (cfft_c64_poly_*): twist[j] = e^{+i pi j / (2n)} through the
        // reference's sincospi64 (src/fft_simd.rs:237-296; j / 2n is exact), untwist[j] = conj(twist[j]) / n (exact, n = 2^k)
        std::vector<cplx> tw(2 * p->n);
        const double inv_n = 1.0 / double(p->n);
        for (uint64_t j = 0; j < p->n; j++) {
            double s, c;
            sincospi64(double(j) / double(2 * p->n), s, c);
            tw[j] = cplx{c, s};
            tw[p->n + j] = cplx{c * inv_n, -s * inv_n};
        }
        CU(cudaMalloc(reinterpret_cast<void **>(&p->d_twist), 2 * p->n * sizeof(cplx)));
        CU(cudaMemcpy(p->d_twist, tw.data(), 2 * p->n * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    return CFFT_OK;
}

// Planar copies w_k[p] (k-major) of the unordered level tables for the register kernel of c64_regs.cu:
// lanes on consecutive p then read consecutive entries.  Same values as h_tw (the reference lays these
// tables out per SIMD width itself, src/unordered.rs:373-385).
cfft_status build_top_planar(cfft_plan *p)
{
    for (int d = 0; d < 2; d++) {
        std::vector<cplx> out;
        StageProgram &pg = p->prog[d];
        for (int i = 0; i < pg.count; i++) {
            Stage &st = pg.st[i];
            st.tw2 = 0;
            if (st.kind != ST_TOP) continue;
            st.tw2 = uint32_t(out.size());
            const uint32_t r = uint32_t(st.radix), m = st.span / r;
            const cplx *src = p->h_tw[d].data() + st.tw_off;
            for (uint32_t k = 1; k < r; k++)
                for (uint32_t q = 0; q < m; q++) out.push_back(src[size_t(r - 1) * q + (k - 1)]);
        }
        if (out.empty()) continue;
        CU(cudaMalloc(reinterpret_cast<void **>(&p->d_top_tw[d]), out.size() * sizeof(cplx)));
        CU(cudaMemcpy(p->d_top_tw[d], out.data(), out.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    return CFFT_OK;
}

const char *variant_name(const cfft_plan *p)
{
    if (p->kind == KIND_F128) return "f128-radix8-tile";
    switch (p->fast_variant) {
    case 1: return "fast-b256-regs";
    case 2: return "fast-b256-column+rows";
    case 3: return "ordered-b256-column+rows-std";
    case 4: return "fast-b256-cluster";
    case 5: return "ordered-b256-regs-std";
    case 6: return "ord16-regs";
    case 8: return "fast-b256-persistent-2pass";
    case 9: return "fast-b256-column+fused-rows";
    default:
        if (p->exact_regs && (p->kind == KIND_UNORDERED || !p->allow_large) && !getenv("CFFT_B200_REGS_NO_SPEC") &&
            regs_spec_supported(p->n, algo_radix(p->algo), algo_is_dit(p->algo), p->kind == KIND_ORDERED ? p->n : p->base_n))
            return "exact-regs-spec"; // compile-time schedule (c64_regs.cu)
        return p->exact_regs ? "exact-regs" : "exact-tile";
    }
}

// Planar copies of the plan's twiddles for c64_fast.cu: per unordered level w_k[p] (k-major,
// (r-1) x m), then the planar half of the base init_wt table.  Same values, different layout
// (the reference itself lays the level tables out per SIMD width, src/unordered.rs:373-385).
cfft_status build_fast_tables(cfft_plan *p)
{
    const bool ordered_large = p->kind == KIND_ORDERED && p->allow_large;
    // the ordered Dif16 plan of size 256 IS the 256-point base FFT of the register kernel
    const bool ordered_256 = p->kind == KIND_ORDERED && !p->allow_large && p->n == 256 && p->algo == CFFT_DIF16;
    if (!ordered_large) {
        if (getenv("CFFT_B200_FORCE_EXACT")) return CFFT_OK;
        // whole-transform Dif16 plans of the other sizes: c64_ord16.cu, straight from the plan's own table
        const bool whole = p->kind == KIND_ORDERED || (p->kind == KIND_UNORDERED && p->base_n == p->n);
        if (whole && ord16_supported(p->n, p->algo)) {
            p->fast_variant = 6;
            p->kernel_name = variant_name(p);
            return CFFT_OK;
        }
        if (!ordered_256 && (p->kind != KIND_UNORDERED || !fast_b256_supported(p->n, p->algo, p->base_n))) return CFFT_OK;
    }
    // levels top-down; forward offsets from prog[0], inverse offsets from prog[1] (stored bottom-up)
    std::vector<Stage> tops_f, tops_i;
    for (int i = 0; i < p->prog[0].count; i++) if (p->prog[0].st[i].kind == ST_TOP) tops_f.push_back(p->prog[0].st[i]);
    for (int i = p->prog[1].count - 1; i >= 0; i--) if (p->prog[1].st[i].kind == ST_TOP) tops_i.push_back(p->prog[1].st[i]);
    for (int d = 0; d < 2; d++) {
        std::vector<cplx> out;
        p->fast_levels.clear();
        for (size_t i = 0; i < tops_f.size(); i++) {
            const Stage &st = d == 0 ? tops_f[i] : tops_i[i];
            const uint32_t r = uint32_t(st.radix), m = st.span / r;
            p->fast_levels.push_back(cfft_plan::FastLevel{st.radix, st.span, uint32_t(out.size())});
            const cplx *src = p->h_tw[d].data() + st.tw_off;
            for (uint32_t k = 1; k < r; k++)
                for (uint32_t q = 0; q < m; q++) out.push_back(src[size_t(r - 1) * q + (k - 1)]);
        }
        p->fast_base_off = uint32_t(out.size());
        // base table: forward at the end of the level tables, inverse at offset 0; planar half first
        const size_t base_off = d == 0 ? p->h_tw[0].size() - 2 * p->base_n : 0;
        for (uint64_t i = 0; i < p->base_n; i++) out.push_back(p->h_tw[d][base_off + i]);
        CU(cudaMalloc(reinterpret_cast<void **>(&p->d_fast_tw[d]), out.size() * sizeof(cplx)));
        CU(cudaMemcpy(p->d_fast_tw[d], out.data(), out.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    }
    if (p->n == 256) {
        p->fast_variant = 1;
        p->kernel_name = variant_name(p);
        return CFFT_OK;
    }
    // n > 8192: group the levels (8, 8, ..., 8, [4|2]) into as few HBM passes as possible, up to three
    // levels (combined radix <= 512) per pass, the levels spread evenly over the passes (inner passes larger).
    const int nl = int(p->fast_levels.size());
    {
        const int ngroups = (nl + 2) / 3, small = nl / ngroups, extra = nl % ngroups;
        int first = 0;
        for (int gi = 0; gi < ngroups; gi++) {
            const int len = small + (gi >= ngroups - extra ? 1 : 0);
            cfft_plan::FastGroup g{{1, 1, 1}, p->fast_levels[size_t(first)].span, first};
            for (int i = 0; i < len; i++) g.radices[i] = p->fast_levels[size_t(first + i)].radix;
            p->fast_groups.push_back(g);
            first += len;
        }
    }
    // Variant 9: peel the largest tail of levels that a fused kernel covers (256 x tail radices <= 4096) and
    // group only the levels above it, three radix-8 levels (512 rows) per column pass at most.
    if (p->n >= 16384 && !ordered_large) {
        int t = 1;
        uint32_t tail = 256u * uint32_t(p->fast_levels[size_t(nl - 1)].radix);
        if (nl >= 2 && tail * uint32_t(p->fast_levels[size_t(nl - 2)].radix) <= 4096) {
            t = 2;
            tail *= uint32_t(p->fast_levels[size_t(nl - 2)].radix);
        }
        p->tail_n = tail;
        p->tail_first_level = nl - t;
        const int k = nl - t; // upper levels, all radix 8
        bool ok = k >= 1;
        for (int i = 0; i < k; i++) ok = ok && p->fast_levels[size_t(i)].radix == 8;
        if (ok) {
            const int ngroups = (k + 2) / 3;
            int first = 0;
            for (int gi = 0; gi < ngroups; gi++) {
                const int len = (k - first + (ngroups - gi) - 1) / (ngroups - gi); // spread evenly, larger groups first
                cfft_plan::FastGroup g{{1, 1, 1}, p->fast_levels[size_t(first)].span, first};
                for (int i = 0; i < len; i++) g.radices[i] = 8;
                p->tail_groups.push_back(g);
                first += len;
            }
        } else {
            p->tail_n = 0;
        }
    }
    // n <= 8192 also has the fused single-kernel variant, the default there (autotune may switch)
    p->fast_variant = ordered_large ? (p->n <= 8192 ? 5 : 3) : (p->n <= 8192 ? 1 : (p->n >= 131072 && p->tail_n ? 9 : 2));
    if (const char *fv = getenv("CFFT_B200_FAST_VARIANT")) { // testing hook: force a variant
        if (ordered_large && atoi(fv) == 3) p->fast_variant = 3;
        if (!ordered_large && atoi(fv) == 2) p->fast_variant = 2;
        if (!ordered_large && atoi(fv) == 4 && (p->n == 8192 || p->n == 16384)) p->fast_variant = 4;
        if (!ordered_large && atoi(fv) == 8 && p->n >= 16384 && p->n <= 65536) p->fast_variant = 8;
        if (!ordered_large && atoi(fv) == 9 && p->tail_n) p->fast_variant = 9;
    }
    p->kernel_name = variant_name(p);
    return CFFT_OK;
}

cfft_status check_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(CFFT_ECUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                    cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(CFFT_EINVAL, "device index out of range");
    return CFFT_OK;
}

// Method::Measure replacement: deterministic per-size choice.  The
// Fourier-domain order is a function of base_n, so it must not depend on a timing race.
void measure_choice(uint64_t n, int *algo, uint64_t *base_n)
{
    *algo = CFFT_DIF16; // radix-16 stages minimise the number of shared-memory exchanges
    if (n <= 256) *base_n = n;                       // as the reference, src/unordered.rs:561-564
    else *base_n = 256;                              // register / column kernels (c64_fast.cu)
}

} // namespace

namespace {

// time fwd+inv of the plan's current variant on a scratch batch: best of 3 after one warm-up
cfft_status time_variant(cfft_plan *p, void *scratch, uint64_t batch, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1,
                         float *ms_out)
{
    float best = 1e30f;
    for (int it = 0; it < 4; it++) {
        CU(cudaEventRecord(e0, st));
        for (int dir = 0; dir < 2; dir++) {
            cudaError_t e;
            if (p->kind == KIND_F128) {
                double *b = static_cast<double *>(scratch);
                const uint64_t pl = batch * p->n;
                e = launch_f128(p, dir == 1, b, b + pl, b + 2 * pl, b + 3 * pl, batch, st);
            } else {
                e = launch_c64(p, dir == 1, static_cast<double2 *>(scratch), batch, st);
            }
            if (e != cudaSuccess) return cuda_fail(e, "autotune launch");
        }
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (it > 0 && ms < best) best = ms;
    }
    *ms_out = best;
    return CFFT_OK;
}

} // namespace

namespace cfft {
cudaError_t launch_c64(const cfft_plan *plan, bool inverse, double2 *data, uint64_t batch, cudaStream_t st)
{
    if (plan->fast_variant == 6) return launch_c64_ord16(plan, inverse, data, batch, st);
    if (plan->fast_variant != 0) return launch_c64_fast_b256(plan, inverse, data, batch, st);
    return launch_c64_exact(plan, inverse, data, batch, st);
}
void count_launch(uint64_t k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
cfft_status set_last_error(cfft_status st, const std::string &msg) { return fail(st, msg); }
} // namespace cfft

extern "C" {

const char *cfft_status_string(cfft_status st)
{
    switch (st) {
    case CFFT_OK: return "ok";
    case CFFT_EINVAL: return "invalid argument";
    case CFFT_ECUDA: return "CUDA error";
    case CFFT_ENOMEM: return "out of memory";
    case CFFT_EUNSUPPORTED: return "unsupported";
    case CFFT_ELENGTH: return "buffer length does not match the plan";
    default: return "unknown status";
    }
}
const char *cfft_last_error(void) { return g_last_error.c_str(); }
uint64_t cfft_launch_count(void) { return g_launches.load(); }
const char *cfft_version(void) { return "cfft_b200 0.1.0 sm_100a"; }

cfft_status cfft_ordered_plan_create(cfft_plan **out, int device, uint64_t n, int method, int algo, int allow_large)
{
    if (!out) return fail(CFFT_EINVAL, "out is null");
    *out = nullptr;
    if (!is_pow2(n)) return fail(CFFT_EINVAL, "n must be a power of two (src/ordered.rs:243)");
    if (ilog2(n) >= 11 && !allow_large) return fail(CFFT_EINVAL, "ordered plans need n <= 2^10 (src/ordered.rs:244)");
    if (ilog2(n) > 20) return fail(CFFT_EINVAL, "ordered plans are limited to n <= 2^20");
    const bool large = ilog2(n) >= 11;
    if (method == CFFT_METHOD_MEASURE) algo = CFFT_DIF16;
    else if (method != CFFT_METHOD_USER) return fail(CFFT_EINVAL, "unknown method");
    if (algo < CFFT_DIF2 || algo > CFFT_DIT16) return fail(CFFT_EINVAL, "unknown FftAlgo");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");

    cfft_plan *p = new (std::nothrow) cfft_plan;
    if (!p) return fail(CFFT_ENOMEM, "plan allocation failed");
    p->kind = KIND_ORDERED;
    p->device = device;
    p->n = n;
    p->algo = algo;
    p->base_n = n;
    p->method = method;
    p->kernel_name = variant_name(p);
    if (large) {
        // Extension (no reference implementation, src/ordered.rs:244): X = DFT(x) in standard order,
        // computed as the unordered plan (Dif16, 256) with the un-permutation fused into the last pass.
        p->allow_large = true;
        p->algo = CFFT_DIF16;
        p->base_n = 256; // internal base size; cfft_plan_algo reports n for ordered plans
        init_unordered_twiddles(n, 256, 16, p->h_tw[0], p->h_tw[1]);
        build_c64_programs(p);
        st = upload_c64(p);
        if (st == CFFT_OK) st = build_fast_tables(p);
        if (st == CFFT_OK && p->fast_variant == 0) st = build_top_planar(p); // only the generic kernels read these copies
        if (st == CFFT_OK && method == CFFT_METHOD_MEASURE && !getenv("CFFT_B200_NO_AUTOTUNE")) st = cfft_plan_autotune(p, 0);
        if (st != CFFT_OK) { cfft_plan_destroy(p); return st; }
        *out = p;
        return CFFT_OK;
    }
    // src/ordered.rs:259-270: zero-initialised 2n tables, filled by init_wt
    p->h_tw[0].assign(2 * n, cplx{0.0, 0.0});
    p->h_tw[1].assign(2 * n, cplx{0.0, 0.0});
    init_wt(size_t(algo_radix(algo)), n, p->h_tw[0].data(), p->h_tw[1].data());
    // the ordered table is the unordered layout with zero levels: [base table (2n)]
    StageProgram &f = p->prog[0], &v = p->prog[1];
    f.count = v.count = 0;
    append_base_stages(f, algo, n, uint32_t(n));
    append_base_stages(v, algo, n, uint32_t(n));
    st = upload_c64(p);
    if (st == CFFT_OK) st = build_fast_tables(p);
    if (st == CFFT_OK && p->fast_variant == 0) st = build_top_planar(p); // only the generic kernels read these copies
    if (st != CFFT_OK) { cfft_plan_destroy(p); return st; }
    *out = p;
    return CFFT_OK;
}

cfft_status cfft_unordered_plan_create(cfft_plan **out, int device, uint64_t n, int method, int base_algo,
                                       uint64_t base_n)
{
    if (!out) return fail(CFFT_EINVAL, "out is null");
    *out = nullptr;
    if (!is_pow2(n)) return fail(CFFT_EINVAL, "n must be a power of two (src/unordered.rs:660)");
    if (n > (uint64_t{1} << 26)) return fail(CFFT_EINVAL, "n above 2^26 is not supported");
    if (method == CFFT_METHOD_MEASURE) measure_choice(n, &base_algo, &base_n);
    else if (method != CFFT_METHOD_USER) return fail(CFFT_EINVAL, "unknown method");
    if (base_algo < CFFT_DIF2 || base_algo > CFFT_DIT16) return fail(CFFT_EINVAL, "unknown FftAlgo");
    // src/unordered.rs:664-669
    if (!is_pow2(base_n)) return fail(CFFT_EINVAL, "base_n must be a power of two");
    if (base_n > n) return fail(CFFT_EINVAL, "base_n must be <= n");
    if (base_n != n && base_n < 32) return fail(CFFT_EINVAL, "base_n must be >= 32 unless it equals n");
    if (ilog2(base_n) > 10) return fail(CFFT_EINVAL, "base_n must be <= 1024");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");

    cfft_plan *p = new (std::nothrow) cfft_plan;
    if (!p) return fail(CFFT_ENOMEM, "plan allocation failed");
    p->kind = KIND_UNORDERED;
    p->device = device;
    p->n = n;
    p->algo = base_algo;
    p->base_n = base_n;
    p->method = method;
    p->kernel_name = variant_name(p);
    init_unordered_twiddles(n, base_n, size_t(algo_radix(base_algo)), p->h_tw[0], p->h_tw[1]);
    build_c64_programs(p);
    st = upload_c64(p);
    if (st == CFFT_OK) st = build_fast_tables(p);
    if (st == CFFT_OK && p->fast_variant == 0) st = build_top_planar(p); // only the generic kernels read these copies
    if (st == CFFT_OK && method == CFFT_METHOD_MEASURE && !getenv("CFFT_B200_NO_AUTOTUNE")) st = cfft_plan_autotune(p, 0);
    if (st != CFFT_OK) { cfft_plan_destroy(p); return st; }
    *out = p;
    return CFFT_OK;
}

cfft_status cfft_f128_plan_create(cfft_plan **out, int device, uint64_t n)
{
    if (!out) return fail(CFFT_EINVAL, "out is null");
    *out = nullptr;
    if (!is_pow2(n)) return fail(CFFT_EINVAL, "n must be a power of two (src/fft128/mod.rs:1865)");
    if (n < 32) return fail(CFFT_EINVAL, "n must be >= 32 (src/fft128/mod.rs:1866)");
    if (n > (uint64_t{1} << 26)) return fail(CFFT_EINVAL, "n above 2^26 is not supported");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");

    cfft_plan *p = new (std::nothrow) cfft_plan;
    if (!p) return fail(CFFT_ENOMEM, "plan allocation failed");
    p->kind = KIND_F128;
    p->device = device;
    p->n = n;
    p->base_n = n;
    p->kernel_name = "f128-radix8-tile";
    for (int i = 0; i < 4; i++) p->h_f128_tw[i].assign(n, 0.0);
    init_negacyclic_twiddles(n, p->h_f128_tw[0].data(), p->h_f128_tw[1].data(), p->h_f128_tw[2].data(),
                             p->h_f128_tw[3].data());
    for (int i = 0; i < 4; i++) {
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p->d_f128_tw[i]), n * sizeof(double));
        if (e == cudaSuccess)
            e = cudaMemcpy(p->d_f128_tw[i], p->h_f128_tw[i].data(), n * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cfft_plan_destroy(p); return cuda_fail(e, "f128 twiddle upload"); }
    }
    {
        std::vector<double> tw4(4 * n);
        for (uint64_t i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) tw4[4 * i + k] = p->h_f128_tw[k][i];
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&p->d_f128_tw4), 4 * n * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpy(p->d_f128_tw4, tw4.data(), 4 * n * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cfft_plan_destroy(p); return cuda_fail(e, "f128 twiddle upload"); }
    }
    *out = p;
    return CFFT_OK;
}

void cfft_plan_destroy(cfft_plan *p)
{
    if (!p) return;
    DeviceGuard guard(p->device);
    for (int d = 0; d < 2; d++) if (p->d_tw[d]) cudaFree(p->d_tw[d]);
    for (int d = 0; d < 2; d++) if (p->d_fast_tw[d]) cudaFree(p->d_fast_tw[d]);
    for (int d = 0; d < 2; d++) if (p->d_top_tw[d]) cudaFree(p->d_top_tw[d]);
    if (p->d_monomial_tw) cudaFree(p->d_monomial_tw);
    if (p->d_twist) cudaFree(p->d_twist);
    for (int i = 0; i < 4; i++) if (p->d_f128_tw[i]) cudaFree(p->d_f128_tw[i]);
    if (p->d_f128_tw4) cudaFree(p->d_f128_tw4);
    delete p;
}

cfft_status cfft_plan_clone(const cfft_plan *p, cfft_plan **out) { return cfft_plan_clone_to_device(p, p ? p->device : 0, out); }

// a replica of the plan -- same transform, same Fourier-domain order, same tuned kernel variant -- with its tables on `device`
cfft_status cfft_plan_clone_to_device(const cfft_plan *p, int device, cfft_plan **out)
{
    if (!p || !out) return fail(CFFT_EINVAL, "null argument");
    cfft_status st;
    switch (p->kind) {
    case KIND_ORDERED: st = cfft_ordered_plan_create(out, device, p->n, CFFT_METHOD_USER, p->algo, p->allow_large); break;
    case KIND_UNORDERED: st = cfft_unordered_plan_create(out, device, p->n, CFFT_METHOD_USER, p->algo, p->base_n); break;
    default: st = cfft_f128_plan_create(out, device, p->n); break;
    }
    if (st == CFFT_OK) { // keep the source plan's tuned variant
        (*out)->method = p->method;
        (*out)->fast_variant = p->fast_variant;
        (*out)->tile_elems = p->tile_elems;
        (*out)->f128_smax = p->f128_smax;
        (*out)->exact_regs = p->exact_regs;
        (*out)->l2_chunk_mb = p->l2_chunk_mb;
        (*out)->l2_streams = p->l2_streams;
        (*out)->kernel_name = p->kernel_name;
        (*out)->tuning_report = p->tuning_report;
    }
    return st;
}

uint64_t cfft_plan_fft_size(const cfft_plan *p) { return p ? p->n : 0; }
int cfft_plan_kind(const cfft_plan *p) { return p ? p->kind : -1; }
int cfft_plan_device(const cfft_plan *p) { return p ? p->device : -1; }
const char *cfft_plan_kernel_name(const cfft_plan *p) { return p ? p->kernel_name.c_str() : ""; }

cfft_status cfft_plan_algo(const cfft_plan *p, int *algo, uint64_t *base_n)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (algo) *algo = p->algo;
    if (base_n) *base_n = (p->kind == KIND_ORDERED) ? p->n : p->base_n;
    return CFFT_OK;
}

cfft_status cfft_plan_scratch_req(const cfft_plan *p, uint64_t *bytes, uint64_t *align)
{
    if (!p) return fail(CFFT_EINVAL, "null plan");
    // ordered: n c64 (src/ordered.rs:320-322); unordered: base_n c64 (src/unordered.rs:798-800);
    // fft128 needs none.  CACHELINE_ALIGN of aligned-vec 0.5 is 128 on x86-64.
    if (bytes) *bytes = (p->kind == KIND_F128) ? 0 : (p->kind == KIND_ORDERED ? p->n : p->base_n) * sizeof(cplx);
    if (align) *align = 128;
    return CFFT_OK;
}

cfft_status cfft_plan_autotune(cfft_plan *p, uint64_t batch_hint)
{
    if (!p) return fail(CFFT_EINVAL, "null plan");
    if (p->n < 2) return CFFT_OK;
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    const uint64_t bytes_per = p->n * (p->kind == KIND_F128 ? 32u : 16u);
    uint64_t batch = batch_hint ? batch_hint : std::max<uint64_t>(1, (uint64_t((p->n >= 16384 || p->fast_variant == 3 || p->fast_variant == 5) ? 512 : 128) << 20) / bytes_per);
    if (batch * bytes_per > (uint64_t{1} << 30)) batch = std::max<uint64_t>(1, (uint64_t{1} << 30) / bytes_per);

    struct Cand { std::string name; int fast_variant; uint32_t tile; uint32_t l2_mb = 0, l2_streams = 1; int smax = 3; bool regs = false; };
    std::vector<Cand> cands;
    if (p->kind == KIND_F128 || p->fast_variant == 0 || p->fast_variant == 6) {
        const char *fam = p->kind == KIND_F128 ? "f128-radix8-tile" : "exact-tile";
        if (p->fast_variant == 6) cands.push_back({"ord16-regs", 6, 0});
        if (p->kind != KIND_F128) cands.push_back({"exact-regs", 0, 0, 0, 1, 3, true});
        if (p->n <= 2048)
            for (uint32_t t : {1024u, 2048u, 4096u})
                if (t >= p->n) cands.push_back({std::string(fam) + "/" + std::to_string(t), p->kind == KIND_F128 ? p->fast_variant : 0, t});
        if (p->kind != KIND_F128 && p->n > 2048) cands.push_back({"exact-tile/4096", 0, 0});
        if (p->kind == KIND_F128 && p->n >= 4096) { // one CTA per SM on 4096-element tiles, or one more HBM pass and two CTAs per SM
            cands.push_back({std::string(fam) + "/4096", p->fast_variant, 4096u});
            cands.push_back({std::string(fam) + "/2048+hbm-pass", p->fast_variant, 2048u});
        }
        if (p->kind == KIND_F128 && p->n <= 2048) // two-stage groups in 80 registers: three CTAs per SM
            cands.push_back({std::string(fam) + "/2048/3-per-SM", p->fast_variant, 2048u, 0, 1, 2});
    } else if (p->fast_variant == 3 || p->fast_variant == 5) {
        if (p->n <= 8192) cands.push_back({"ordered-b256-regs-std", 5, 0});
        cands.push_back({"ordered-b256-column+rows-std", 3, 0});
        cands.push_back({"ordered-b256-column+rows-std/L2-8MBx4", 3, 0, 8, 4}); // out of place: chunk + workspace share L2
        cands.push_back({"ordered-b256-column+rows-std/L2-16MBx4", 3, 0, 16, 4});
        cands.push_back({"ordered-b256-column+rows-std/L2-32MBx2", 3, 0, 32, 2});
    } else if (p->fast_variant == 1 || p->fast_variant == 2 || p->fast_variant == 4 || p->fast_variant == 8 || p->fast_variant == 9) {
        if (p->tail_n) {
            cands.push_back({"fast-b256-column+fused-rows", 9, 0});
            cands.push_back({"fast-b256-column+fused-rows/L2-16MBx4", 9, 0, 16, 4});
            cands.push_back({"fast-b256-column+fused-rows/L2-32MBx2", 9, 0, 32, 2});
        }
        // the persistent two-phase kernel spin-waits on other CTAs: opt-in, never picked silently
        if (p->n >= 16384 && p->n <= 65536 && getenv("CFFT_B200_ALLOW_PERSISTENT")) cands.push_back({"fast-b256-persistent-2pass", 8, 0});
        if (p->n > 256 && p->n <= 8192) cands.push_back({"fast-b256-regs", 1, 0});
        if (p->n > 256 && p->n <= 16384) cands.push_back({"fast-b256-column+rows", 2, 0});
        if (p->n == 8192 || p->n == 16384) cands.push_back({"fast-b256-cluster", 4, 0});
        if (p->n >= 16384) {
            if (p->n > 16384) cands.push_back({"fast-b256-column+rows", 2, 0});
            cands.push_back({"fast-b256-column+rows/L2-16MBx4", 2, 0, 16, 4});
            cands.push_back({"fast-b256-column+rows/L2-24MBx3", 2, 0, 24, 3});
            cands.push_back({"fast-b256-column+rows/L2-32MBx2", 2, 0, 32, 2});
        }
    }
    if (cands.size() < 2) {
        p->tuning_report = p->kernel_name + ": only variant\n";
        return CFFT_OK;
    }
    void *scratch = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cfft_status rc = CFFT_OK;
    auto cleanup = [&] {
        if (scratch) cudaFree(scratch);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
    };
    cudaError_t ce = cudaMalloc(&scratch, batch * bytes_per);
    if (ce == cudaSuccess) ce = cudaMemset(scratch, 0, batch * bytes_per);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce != cudaSuccess) { cleanup(); return cuda_fail(ce, "autotune setup"); }

    const int keep_variant = p->fast_variant;
    const uint32_t keep_tile = p->tile_elems, keep_mb = p->l2_chunk_mb, keep_st = p->l2_streams;
    const int keep_smax = p->f128_smax;
    const bool keep_regs = p->exact_regs;
    std::string report;
    float best_ms = 1e30f;
    size_t best = 0;
    for (size_t i = 0; i < cands.size() && rc == CFFT_OK; i++) {
        p->fast_variant = cands[i].fast_variant;
        p->tile_elems = cands[i].tile;
        p->f128_smax = cands[i].smax;
        p->exact_regs = cands[i].regs;
        p->l2_chunk_mb = cands[i].l2_mb;
        p->l2_streams = cands[i].l2_streams;
        float ms = 0;
        rc = time_variant(p, scratch, batch, st, e0, e1, &ms);
        if (rc != CFFT_OK) break;
        char line[160];
        snprintf(line, sizeof line, "%s: %.4f ms (fwd+inv, batch %llu)\n", cands[i].name.c_str(), ms,
                 static_cast<unsigned long long>(batch));
        report += line;
        if (ms < best_ms) { best_ms = ms; best = i; }
    }
    cleanup();
    {   // give back what the timed candidates left in the stream-ordered workspace pool (a Measure plan must
        // not keep hundreds of MiB of device memory for the life of the process)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceSynchronize() == cudaSuccess && workspace_pool(p->device, &pool) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    if (rc != CFFT_OK) {
        p->fast_variant = keep_variant;
        p->tile_elems = keep_tile;
        p->f128_smax = keep_smax;
        p->exact_regs = keep_regs;
        p->l2_chunk_mb = keep_mb;
        p->l2_streams = keep_st;
        return rc;
    }
    p->fast_variant = cands[best].fast_variant;
    p->tile_elems = cands[best].tile;
    p->f128_smax = cands[best].smax;
    p->exact_regs = cands[best].regs;
    p->l2_chunk_mb = cands[best].l2_mb;
    p->l2_streams = cands[best].l2_streams;
    p->kernel_name = variant_name(p);
    if (p->l2_chunk_mb) p->kernel_name += "/L2-chunked";
    p->tuning_report = report + "selected: " + cands[best].name + "\n";
    return CFFT_OK;
}

uint64_t cfft_plan_tuning_report(const cfft_plan *p, char *buf, uint64_t buf_len)
{
    if (!p || !buf || buf_len == 0) return 0;
    const uint64_t nbytes = std::min<uint64_t>(buf_len - 1, p->tuning_report.size());
    memcpy(buf, p->tuning_report.data(), nbytes);
    buf[nbytes] = 0;
    return nbytes;
}

cfft_status cfft_plan_copy_twiddles(const cfft_plan *p, int which, void *host_out, uint64_t bytes)
{
    if (!p || !host_out) return fail(CFFT_EINVAL, "null argument");
    DeviceGuard guard(p->device);
    if (p->kind == KIND_F128) {
        if (which < 0 || which > 3 || bytes != p->n * sizeof(double)) return fail(CFFT_EINVAL, "bad table / size");
        CU(cudaMemcpy(host_out, p->d_f128_tw[which], bytes, cudaMemcpyDeviceToHost));
        return CFFT_OK;
    }
    if (which < 0 || which > 1 || bytes != p->h_tw[which].size() * sizeof(cplx)) return fail(CFFT_EINVAL, "bad table / size");
    CU(cudaMemcpy(host_out, p->d_tw[which], bytes, cudaMemcpyDeviceToHost));
    return CFFT_OK;
}

// ---- device entry points -----------------------------------------------------------------

static cfft_status run_c64(const cfft_plan *p, bool inverse, void *dev_buf, uint64_t batch, void *stream)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (!dev_buf && batch) return fail(CFFT_EINVAL, "null buffer");
    if (reinterpret_cast<uintptr_t>(dev_buf) & 15) return fail(CFFT_EINVAL, "device buffer must be 16-byte aligned (128-bit accesses)");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64(p, inverse, static_cast<double2 *>(dev_buf), batch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, inverse ? "c64 inv launch" : "c64 fwd launch");
    return CFFT_OK;
}

cfft_status cfft_c64_fwd(const cfft_plan *p, void *dev_buf, uint64_t batch, void *stream)
{
    return run_c64(p, false, dev_buf, batch, stream);
}
cfft_status cfft_c64_inv(const cfft_plan *p, void *dev_buf, uint64_t batch, void *stream)
{
    return run_c64(p, true, dev_buf, batch, stream);
}

// Rows row_stride >= n elements apart (the stride_elems of the ABI proposal -- a caller whose polynomials sit
// inside larger records, e.g. one polynomial of every GLWE ciphertext).  Plans served by one fused register kernel take the
// stride in the kernel's row accessor; every other plan packs <= 256 MiB of rows at a time into the stream-ordered
// workspace, transforms them there and copies them back (two extra passes over the data: correct, not fast).
static cfft_status run_c64_strided(const cfft_plan *p, bool inverse, void *dev_buf, uint64_t row_stride, uint64_t batch, void *stream)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (row_stride == p->n || batch <= 1) return run_c64(p, inverse, dev_buf, batch, stream);
    if (row_stride < p->n) return fail(CFFT_EINVAL, "row stride smaller than the transform size (rows would overlap)");
    if (!dev_buf) return fail(CFFT_EINVAL, "null buffer");
    if (reinterpret_cast<uintptr_t>(dev_buf) & 15) return fail(CFFT_EINVAL, "device buffer must be 16-byte aligned (128-bit accesses)");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double2 *data = static_cast<double2 *>(dev_buf);
    if (fast_b256_strided_available(p)) {
        cudaError_t e = launch_c64_fast_b256_strided(p, inverse, data, row_stride, batch, st);
        if (e != cudaSuccess) return cuda_fail(e, "strided c64 launch");
        return CFFT_OK;
    }
    const uint64_t row_bytes = p->n * sizeof(double2), pitch = row_stride * sizeof(double2);
    uint64_t chunk_rows = std::max<uint64_t>(1, (uint64_t{256} << 20) / row_bytes);
    chunk_rows = std::min(chunk_rows, batch);
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(p->device, &pool);
    if (e != cudaSuccess) return cuda_fail(e, "workspace pool");
    double2 *ws = nullptr;
    if ((e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), chunk_rows * row_bytes, pool, st)) != cudaSuccess)
        return cuda_fail(e, "workspace allocation");
    for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
        const uint64_t rows = std::min(chunk_rows, batch - r0);
        double2 *src = data + r0 * row_stride;
        e = cudaMemcpy2DAsync(ws, row_bytes, src, pitch, row_bytes, rows, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = launch_c64(p, inverse, ws, rows, st);
        if (e == cudaSuccess) e = cudaMemcpy2DAsync(src, pitch, ws, row_bytes, row_bytes, rows, cudaMemcpyDeviceToDevice, st);
    }
    const cudaError_t e2 = cudaFreeAsync(ws, st);
    if (e != cudaSuccess || e2 != cudaSuccess) return cuda_fail(e != cudaSuccess ? e : e2, "strided c64 transform");
    return CFFT_OK;
}

cfft_status cfft_c64_fwd_strided(const cfft_plan *p, void *dev_buf, uint64_t row_stride, uint64_t batch, void *stream)
{
    return run_c64_strided(p, false, dev_buf, row_stride, batch, stream);
}
cfft_status cfft_c64_inv_strided(const cfft_plan *p, void *dev_buf, uint64_t row_stride, uint64_t batch, void *stream)
{
    return run_c64_strided(p, true, dev_buf, row_stride, batch, stream);
}

static cfft_status run_f128(const cfft_plan *p, bool inverse, double *re0, double *re1, double *im0, double *im1,
                            uint64_t batch, void *stream)
{
    if (!p || p->kind != KIND_F128) return fail(CFFT_EINVAL, "not an fft128 plan");
    if (batch && (!re0 || !re1 || !im0 || !im1)) return fail(CFFT_EINVAL, "null buffer");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_f128(p, inverse, re0, re1, im0, im1, batch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, inverse ? "f128 inv launch" : "f128 fwd launch");
    return CFFT_OK;
}

cfft_status cfft_f128_fwd(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t batch,
                          void *stream)
{
    return run_f128(p, false, re0, re1, im0, im1, batch, stream);
}
cfft_status cfft_f128_inv(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t batch,
                          void *stream)
{
    return run_f128(p, true, re0, re1, im0, im1, batch, stream);
}

// fft128 with rows row_stride >= n doubles apart in each of the four planes: packed through the workspace like the
// generic c64 path above (the fft128 kernels are FP64-bound, the two copies cost ~15 % at n = 2048)
static cfft_status run_f128_strided(const cfft_plan *p, bool inverse, double *re0, double *re1, double *im0, double *im1,
                                    uint64_t row_stride, uint64_t batch, void *stream)
{
    if (!p || p->kind != KIND_F128) return fail(CFFT_EINVAL, "not an fft128 plan");
    if (row_stride == p->n || batch <= 1) return run_f128(p, inverse, re0, re1, im0, im1, batch, stream);
    if (row_stride < p->n) return fail(CFFT_EINVAL, "row stride smaller than the transform size (rows would overlap)");
    if (!re0 || !re1 || !im0 || !im1) return fail(CFFT_EINVAL, "null buffer");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint64_t row_bytes = p->n * sizeof(double), pitch = row_stride * sizeof(double);
    uint64_t chunk_rows = std::max<uint64_t>(1, (uint64_t{64} << 20) / row_bytes); // 4 planes x 64 MiB
    chunk_rows = std::min(chunk_rows, batch);
    cudaMemPool_t pool = nullptr;
    cudaError_t e = workspace_pool(p->device, &pool);
    if (e != cudaSuccess) return cuda_fail(e, "workspace pool");
    double *ws = nullptr;
    if ((e = cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), 4 * chunk_rows * row_bytes, pool, st)) != cudaSuccess)
        return cuda_fail(e, "workspace allocation");
    double *planes[4] = {re0, re1, im0, im1};
    for (uint64_t r0 = 0; r0 < batch && e == cudaSuccess; r0 += chunk_rows) {
        const uint64_t rows = std::min(chunk_rows, batch - r0);
        double *w[4];
        for (int i = 0; i < 4 && e == cudaSuccess; i++) {
            w[i] = ws + uint64_t(i) * chunk_rows * p->n;
            e = cudaMemcpy2DAsync(w[i], row_bytes, planes[i] + r0 * row_stride, pitch, row_bytes, rows, cudaMemcpyDeviceToDevice, st);
        }
        if (e == cudaSuccess) e = launch_f128(p, inverse, w[0], w[1], w[2], w[3], rows, st);
        for (int i = 0; i < 4 && e == cudaSuccess; i++)
            e = cudaMemcpy2DAsync(planes[i] + r0 * row_stride, pitch, w[i], row_bytes, row_bytes, rows, cudaMemcpyDeviceToDevice, st);
    }
    const cudaError_t e2 = cudaFreeAsync(ws, st);
    if (e != cudaSuccess || e2 != cudaSuccess) return cuda_fail(e != cudaSuccess ? e : e2, "strided fft128 transform");
    return CFFT_OK;
}
cfft_status cfft_f128_fwd_strided(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t row_stride,
                                  uint64_t batch, void *stream)
{
    return run_f128_strided(p, false, re0, re1, im0, im1, row_stride, batch, stream);
}
cfft_status cfft_f128_inv_strided(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t row_stride,
                                  uint64_t batch, void *stream)
{
    return run_f128_strided(p, true, re0, re1, im0, im1, row_stride, batch, stream);
}

cfft_status cfft_f128_binary_op(int device, int op, const double *a_hi, const double *a_lo, const double *b_hi,
                                const double *b_lo, double *out_hi, double *out_lo, uint64_t len, void *stream)
{
    if (op < CFFT_F128_ADD || op > CFFT_F128_DIV_F64_F64) return fail(CFFT_EINVAL, "unknown f128 operator");
    // which operands are f64 (their lo plane is ignored and may be NULL): a for 9, 12 and 13..16, b for 7, 8, 10, 11 and 13..16
    const bool a_f64 = op == CFFT_F128_SUB_F64_F128 || op == CFFT_F128_DIV_F64_F128 || op >= CFFT_F128_ADD_F64_F64;
    const bool b_f64 = op == CFFT_F128_ADD_F128_F64 || op == CFFT_F128_SUB_F128_F64 || op == CFFT_F128_MUL_F128_F64 ||
                       op == CFFT_F128_DIV_F128_F64 || op >= CFFT_F128_ADD_F64_F64;
    if (len && (!a_hi || (!a_lo && !a_f64) || !b_hi || (!b_lo && !b_f64) || !out_hi || !out_lo)) return fail(CFFT_EINVAL, "null buffer");
    if (a_f64) a_lo = nullptr;
    if (b_f64) b_lo = nullptr;
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_f128_binary(op, a_hi, a_lo, b_hi, b_lo, out_hi, out_lo, len, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "f128 binary op launch");
    return CFFT_OK;
}

cfft_status cfft_f128_unary_op(int device, int op, const double *a_hi, const double *a_lo, double *out_hi, double *out_lo,
                               double *out2_hi, double *out2_lo, uint64_t len, void *stream)
{
    if (op < CFFT_F128_SQR || op > CFFT_F128_IS_NAN) return fail(CFFT_EINVAL, "unknown f128 unary operator");
    if (len && (!a_hi || !a_lo || !out_hi || !out_lo)) return fail(CFFT_EINVAL, "null buffer");
    if (len && op == CFFT_F128_SINCOSPI && (!out2_hi || !out2_lo)) return fail(CFFT_EINVAL, "sincospi needs the second output (cos)");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaStream_t cs = static_cast<cudaStream_t>(stream);
    unsigned int *bad = nullptr;
    if (op == CFFT_F128_SINCOSPI) {
        CU(cudaMalloc(reinterpret_cast<void **>(&bad), sizeof(unsigned int)));
        cudaError_t e0 = cudaMemsetAsync(bad, 0, sizeof(unsigned int), cs);
        if (e0 != cudaSuccess) { cudaFree(bad); return cuda_fail(e0, "f128 sincospi setup"); }
    }
    cudaError_t e = launch_f128_unary(op, a_hi, a_lo, out_hi, out_lo, out2_hi, out2_lo, len, bad, cs);
    if (op == CFFT_F128_SINCOSPI) {
        // the reference panics on inputs outside [-1, 1] (f128_ops.rs:536-538): this one entry point is synchronous so
        // that the same precondition comes back as a status instead of silently as NaNs
        unsigned int h = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h, bad, sizeof(h), cudaMemcpyDeviceToHost, cs);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
        cudaFree(bad);
        if (e == cudaSuccess && h) return fail(CFFT_EINVAL, "only inputs in [-1, 1] are currently supported (f128_ops.rs:537)");
    }
    if (e != cudaSuccess) return cuda_fail(e, "f128 unary op launch");
    return CFFT_OK;
}

cfft_status cfft_f128_compare(int device, const double *a_hi, const double *a_lo, const double *b_hi, const double *b_lo, int8_t *out,
                              uint64_t len, void *stream)
{
    if (len && (!a_hi || !a_lo || !b_hi || !out)) return fail(CFFT_EINVAL, "null buffer");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_f128_compare(a_hi, a_lo, b_hi, b_lo, reinterpret_cast<signed char *>(out), len, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "f128 compare launch");
    return CFFT_OK;
}

cfft_status cfft_f128_cplx_mul_scale(int device, double *l_re0, double *l_re1, double *l_im0, double *l_im1,
                                     const double *r_re0, const double *r_re1, const double *r_im0, const double *r_im1,
                                     double factor, uint64_t len, void *stream)
{
    if (len && (!l_re0 || !l_re1 || !l_im0 || !l_im1 || !r_re0 || !r_re1 || !r_im0 || !r_im1))
        return fail(CFFT_EINVAL, "null buffer");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_f128_cplx_mul_scale(l_re0, l_re1, l_im0, l_im1, r_re0, r_re1, r_im0, r_im1, factor, len,
                                               static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "f128 pointwise product launch");
    return CFFT_OK;
}

cfft_status cfft_f128_fwd_mul_inv(const cfft_plan *p, double *l_re0, double *l_re1, double *l_im0, double *l_im1,
                                  const double *r_re0, const double *r_re1, const double *r_im0, const double *r_im1,
                                  uint64_t rhs_row_stride, double factor, uint64_t batch, void *stream)
{
    if (!p || p->kind != KIND_F128) return fail(CFFT_EINVAL, "not an fft128 plan");
    if (batch && (!l_re0 || !l_re1 || !l_im0 || !l_im1 || !r_re0 || !r_re1 || !r_im0 || !r_im1)) return fail(CFFT_EINVAL, "null buffer");
    if (rhs_row_stride != 0 && rhs_row_stride != p->n) return fail(CFFT_EINVAL, "rhs_row_stride must be 0 (rhs shared by every row) or n");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_f128_fwd_mul_inv(p, l_re0, l_re1, l_im0, l_im1, r_re0, r_re1, r_im0, r_im1, rhs_row_stride == 0, factor,
                                            batch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "f128 fwd-mul-inv launch");
    return CFFT_OK;
}

static cfft_status run_pointwise(int device, void *acc, void *a, const void *b, uint64_t len, void *stream)
{
    if (len && (!a || !b)) return fail(CFFT_EINVAL, "null buffer");
    if ((reinterpret_cast<uintptr_t>(acc) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15)
        return fail(CFFT_EINVAL, "device buffers must be 16-byte aligned");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_pointwise(static_cast<double2 *>(acc), static_cast<double2 *>(a), static_cast<const double2 *>(b), len,
                                         static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 pointwise product launch");
    return CFFT_OK;
}

cfft_status cfft_c64_mul_assign(int device, void *lhs_dev, const void *rhs_dev, uint64_t len, void *stream)
{
    return run_pointwise(device, nullptr, lhs_dev, rhs_dev, len, stream);
}

cfft_status cfft_c64_mul_add_assign(int device, void *acc_dev, const void *a_dev, const void *b_dev, uint64_t len, void *stream)
{
    if (len && !acc_dev) return fail(CFFT_EINVAL, "null buffer");
    return run_pointwise(device, acc_dev, const_cast<void *>(a_dev), b_dev, len, stream);
}

cfft_status cfft_c64_fwd_mul_inv(const cfft_plan *p, const void *a_dev, uint64_t k_terms, const void *b_dev,
                                 uint64_t b_row_stride, void *out_dev, uint64_t batch, void *stream)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (k_terms == 0) return fail(CFFT_EINVAL, "k_terms must be >= 1");
    if (batch && (!a_dev || !b_dev || !out_dev)) return fail(CFFT_EINVAL, "null buffer");
    if ((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(b_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15)
        return fail(CFFT_EINVAL, "device buffers must be 16-byte aligned (128-bit accesses)");
    if (b_row_stride != 0 && b_row_stride < k_terms * p->n) return fail(CFFT_EINVAL, "b_row_stride must be 0 (b shared by every row) or >= k_terms * n");
    {   // the kernels read a and b through __restrict__ pointers while other rows of out are being written: only the exact
        // in-place case (out == a, one term) is allowed, any other intersection of the extents is rejected
        const uint64_t row = p->n * sizeof(cplx);
        const uint64_t a_bytes = batch * k_terms * row, out_bytes = batch * row;
        const uint64_t b_bytes = batch == 0 ? 0 : (b_row_stride == 0 ? k_terms * row : ((batch - 1) * b_row_stride + k_terms * p->n) * sizeof(cplx));
        if (out_dev == a_dev) {
            if (k_terms != 1) return fail(CFFT_EINVAL, "out may alias a only when k_terms == 1");
        } else if (ranges_overlap(out_dev, out_bytes, a_dev, a_bytes)) {
            return fail(CFFT_EINVAL, "out overlaps a without being the same buffer");
        }
        if (ranges_overlap(out_dev, out_bytes, b_dev, b_bytes)) return fail(CFFT_EINVAL, "out must not overlap b");
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_fwd_mul_inv(p, static_cast<const double2 *>(a_dev), k_terms, static_cast<const double2 *>(b_dev),
                                           b_row_stride, static_cast<double2 *>(out_dev), batch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 fwd-mul-inv launch");
    return CFFT_OK;
}

cfft_status cfft_c64_fwd_mul_inv_multi(const cfft_plan *p, const void *a_dev, uint64_t k_terms, const void *b_dev, uint64_t b_row_stride,
                                       uint64_t n_out, void *out_dev, uint64_t batch, void *stream)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (k_terms == 0) return fail(CFFT_EINVAL, "k_terms must be >= 1");
    if (n_out == 0 || n_out > 64) return fail(CFFT_EINVAL, "n_out must be 1 .. 64");
    if (batch && (!a_dev || !b_dev || !out_dev)) return fail(CFFT_EINVAL, "null buffer");
    if ((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(b_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15)
        return fail(CFFT_EINVAL, "device buffers must be 16-byte aligned (128-bit accesses)");
    if (b_row_stride != 0 && b_row_stride < k_terms * n_out * p->n) return fail(CFFT_EINVAL, "b_row_stride must be 0 (b shared by every row) or >= k_terms * n_out * n");
    {   // out is written while a and b are still being read by other rows: no overlap at all (there is no in-place form)
        const uint64_t row = p->n * sizeof(cplx);
        const uint64_t a_bytes = batch * k_terms * row, out_bytes = batch * n_out * row;
        const uint64_t b_bytes = batch == 0 ? 0 : (b_row_stride == 0 ? k_terms * n_out * row : ((batch - 1) * b_row_stride + k_terms * n_out * p->n) * sizeof(cplx));
        if (n_out == 1 && out_dev == a_dev) {
            if (k_terms != 1) return fail(CFFT_EINVAL, "out may alias a only when k_terms == 1 and n_out == 1");
        } else if (ranges_overlap(out_dev, out_bytes, a_dev, a_bytes)) {
            return fail(CFFT_EINVAL, "out overlaps a");
        }
        if (ranges_overlap(out_dev, out_bytes, b_dev, b_bytes)) return fail(CFFT_EINVAL, "out must not overlap b");
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_fwd_mul_inv_multi(p, static_cast<const double2 *>(a_dev), k_terms, static_cast<const double2 *>(b_dev), b_row_stride,
                                                 n_out, static_cast<double2 *>(out_dev), batch, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 fwd-mul-inv (several outputs) launch");
    return CFFT_OK;
}

int cfft_plan_has_fused_mul2_kernel(const cfft_plan *p) { return (p && p->kind != KIND_F128 && fused_mul2_kernel_available(p)) ? 1 : 0; }

cfft_status cfft_c64_fwd_mul_add(const cfft_plan *p, const void *a_dev, uint64_t a_row_stride, const void *b_dev,
                                 uint64_t b_row_stride, void *acc_dev, int accumulate, uint64_t batch, void *stream)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (batch && (!a_dev || !b_dev || !acc_dev)) return fail(CFFT_EINVAL, "null buffer");
    if ((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(b_dev) | reinterpret_cast<uintptr_t>(acc_dev)) & 15)
        return fail(CFFT_EINVAL, "device buffers must be 16-byte aligned (128-bit accesses)");
    if (a_row_stride == 0 || a_row_stride % p->n) return fail(CFFT_EINVAL, "a_row_stride must be a positive multiple of n");
    if (b_row_stride != 0 && b_row_stride < p->n) return fail(CFFT_EINVAL, "b_row_stride must be 0 (b shared by every row) or >= n");
    {
        const uint64_t row = p->n * sizeof(cplx);
        const uint64_t a_bytes = batch == 0 ? 0 : ((batch - 1) * a_row_stride + p->n) * sizeof(cplx);
        const uint64_t b_bytes = batch == 0 ? 0 : (b_row_stride == 0 ? row : ((batch - 1) * b_row_stride + p->n) * sizeof(cplx));
        if (ranges_overlap(acc_dev, batch * row, a_dev, a_bytes) || ranges_overlap(acc_dev, batch * row, b_dev, b_bytes))
            return fail(CFFT_EINVAL, "acc must not overlap a or b");
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_fwd_mul_add(p, static_cast<const double2 *>(a_dev), a_row_stride / p->n, static_cast<const double2 *>(b_dev),
                                           b_row_stride, static_cast<double2 *>(acc_dev), accumulate != 0, batch,
                                           static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 fwd-mul-add launch");
    return CFFT_OK;
}

// ---- integer polynomials <-> Fourier domain -------------------------------------------------------------------------
static cfft_status poly_args(const cfft_plan *p, const void *poly, const void *fourier, uint64_t batch, uint32_t flags, bool allow_acc)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (flags & ~3u) return fail(CFFT_EINVAL, "unknown flag bits");
    if (!allow_acc && (flags & 2u)) return fail(CFFT_EINVAL, "CFFT_POLY_ACCUMULATE applies to polynomial outputs only");
    if (batch && (!poly || !fourier)) return fail(CFFT_EINVAL, "null buffer");
    if (reinterpret_cast<uintptr_t>(fourier) & 15) return fail(CFFT_EINVAL, "Fourier-domain buffers must be 16-byte aligned (128-bit accesses)");
    if (reinterpret_cast<uintptr_t>(poly) & 7) return fail(CFFT_EINVAL, "polynomials must be 8-byte aligned");
    return CFFT_OK;
}

cfft_status cfft_c64_poly_fwd(const cfft_plan *p, const int64_t *poly_dev, void *fourier_dev, uint64_t batch, uint32_t flags, void *stream)
{
    cfft_status st = poly_args(p, poly_dev, fourier_dev, batch, flags, false);
    if (st != CFFT_OK) return st;
    const uint64_t bytes = batch * p->n * sizeof(cplx);
    if (ranges_overlap(poly_dev, bytes, fourier_dev, bytes)) return fail(CFFT_EINVAL, "poly and fourier must not overlap");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_poly_fwd(p, reinterpret_cast<const long long *>(poly_dev), static_cast<double2 *>(fourier_dev), batch, flags,
                                        static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 poly fwd launch");
    return CFFT_OK;
}

cfft_status cfft_c64_poly_inv(const cfft_plan *p, const void *fourier_dev, int64_t *poly_dev, uint64_t batch, uint32_t flags, void *stream)
{
    cfft_status st = poly_args(p, poly_dev, fourier_dev, batch, flags, true);
    if (st != CFFT_OK) return st;
    const uint64_t bytes = batch * p->n * sizeof(cplx);
    if (ranges_overlap(poly_dev, bytes, fourier_dev, bytes)) return fail(CFFT_EINVAL, "poly and fourier must not overlap");
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_poly_inv(p, static_cast<const double2 *>(fourier_dev), reinterpret_cast<long long *>(poly_dev), batch, flags,
                                        static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 poly inv launch");
    return CFFT_OK;
}

cfft_status cfft_c64_poly_mul(const cfft_plan *p, const int64_t *a_dev, uint64_t k_terms, const void *b_dev, uint64_t b_row_stride,
                              int64_t *out_dev, uint64_t batch, uint32_t flags, void *stream)
{
    cfft_status st = poly_args(p, a_dev, b_dev, batch, flags, true);
    if (st != CFFT_OK) return st;
    if (k_terms == 0) return fail(CFFT_EINVAL, "k_terms must be >= 1");
    if (batch && !out_dev) return fail(CFFT_EINVAL, "null buffer");
    if (reinterpret_cast<uintptr_t>(out_dev) & 7) return fail(CFFT_EINVAL, "polynomials must be 8-byte aligned");
    if (b_row_stride != 0 && b_row_stride < k_terms * p->n) return fail(CFFT_EINVAL, "b_row_stride must be 0 (b shared by every row) or >= k_terms * n");
    {
        const uint64_t row = p->n * sizeof(cplx); // a polynomial of 2n coefficients is as large as n c64
        const uint64_t b_bytes = batch == 0 ? 0 : (b_row_stride == 0 ? k_terms * row : ((batch - 1) * b_row_stride + k_terms * p->n) * sizeof(cplx));
        if (out_dev == a_dev) {
            if (k_terms != 1 || (flags & 2u)) return fail(CFFT_EINVAL, "out may alias a only when k_terms == 1 and not accumulating");
        } else if (ranges_overlap(out_dev, batch * row, a_dev, batch * k_terms * row)) {
            return fail(CFFT_EINVAL, "out overlaps a without being the same buffer");
        }
        if (ranges_overlap(out_dev, batch * row, b_dev, b_bytes)) return fail(CFFT_EINVAL, "out must not overlap b");
    }
    DeviceGuard guard(p->device);
    if (!guard.ok) return fail(CFFT_ECUDA, "cudaSetDevice failed");
    cudaError_t e = launch_c64_poly_mul(p, reinterpret_cast<const long long *>(a_dev), k_terms, static_cast<const double2 *>(b_dev), b_row_stride,
                                        reinterpret_cast<long long *>(out_dev), batch, flags, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "c64 poly mul launch");
    return CFFT_OK;
}

int cfft_plan_has_fused_poly_kernel(const cfft_plan *p, uint64_t k_terms)
{
    return (p && p->kind != KIND_F128 && poly_fused_available(p, k_terms)) ? 1 : 0;
}

cfft_status cfft_plan_copy_twist(const cfft_plan *p, void *host_out, uint64_t bytes)
{
    if (!p || p->kind == KIND_F128 || !host_out) return fail(CFFT_EINVAL, "not a c64 plan / null out");
    if (bytes != 2 * p->n * sizeof(cplx)) return fail(CFFT_EINVAL, "twist tables hold 2 n c64");
    DeviceGuard guard(p->device);
    CU(cudaMemcpy(host_out, p->d_twist, bytes, cudaMemcpyDeviceToHost));
    return CFFT_OK;
}

cfft_status cfft_twopass_timeouts(int device, uint32_t *out)
{
    if (!out) return fail(CFFT_EINVAL, "null out");
    cfft_status st = check_device(device);
    if (st != CFFT_OK) return st;
    DeviceGuard guard(device);
    unsigned int v = 0;
    CU(twopass_timeouts(&v));
    *out = v;
    return CFFT_OK;
}

int cfft_plan_has_fused_mul_kernel(const cfft_plan *p)
{
    if (!p) return 0;
    return (p->kind == KIND_F128 ? f128_fused_mul_kernel_available(p) : fused_mul_kernel_available(p)) ? 1 : 0;
}

cfft_status cfft_unordered_fwd_monomial(const cfft_plan *p, uint64_t degree, void *dev_buf, void *stream)
{
    if (!p || p->kind != KIND_UNORDERED) return fail(CFFT_EINVAL, "not an unordered plan");
    if (degree >= p->n) return fail(CFFT_EINVAL, "degree must be < n (src/unordered.rs:859)");
    if (!dev_buf) return fail(CFFT_EINVAL, "null buffer");
    DeviceGuard guard(p->device);
    cudaError_t e = launch_monomial(p, degree, static_cast<double2 *>(dev_buf), static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "fwd_monomial launch");
    return CFFT_OK;
}

cfft_status cfft_unordered_permutation(const cfft_plan *p, uint64_t *out)
{
    if (!p || p->kind == KIND_F128 || !out) return fail(CFFT_EINVAL, "not a c64 plan / null out");
    const unsigned nb = ilog2(p->n), bb = ilog2(p->kind == KIND_ORDERED ? p->n : p->base_n);
    for (uint64_t i = 0; i < p->n; i++) out[i] = bit_rev_twice(nb, bb, i);
    return CFFT_OK;
}

static cfft_status run_permute(const cfft_plan *p, bool to_std, const void *src, void *dst, uint64_t batch, void *stream)
{
    if (!p || p->kind == KIND_F128) return fail(CFFT_EINVAL, "not a c64 plan");
    if (batch && (!src || !dst)) return fail(CFFT_EINVAL, "null buffer");
    if (ranges_overlap(src, batch * p->n * sizeof(cplx), dst, batch * p->n * sizeof(cplx))) return fail(CFFT_EINVAL, "src and dst must not overlap");
    DeviceGuard guard(p->device);
    cudaError_t e = launch_permute(p, to_std, static_cast<const double2 *>(src), static_cast<double2 *>(dst), batch,
                                   static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "permute launch");
    return CFFT_OK;
}
cfft_status cfft_unordered_to_standard(const cfft_plan *p, const void *src, void *dst, uint64_t batch, void *stream)
{
    return run_permute(p, true, src, dst, batch, stream);
}
cfft_status cfft_unordered_from_standard(const cfft_plan *p, const void *src, void *dst, uint64_t batch, void *stream)
{
    return run_permute(p, false, src, dst, batch, stream);
}

cfft_status cfft_unordered_to_standard_host(const cfft_plan *p, const void *src, void *dst)
{
    if (!p || p->kind == KIND_F128 || !src || !dst) return fail(CFFT_EINVAL, "bad argument");
    const unsigned nb = ilog2(p->n), bb = p->kind == KIND_ORDERED ? nb : ilog2(p->base_n);
    const cplx *s = static_cast<const cplx *>(src);
    cplx *d = static_cast<cplx *>(dst);
    for (uint64_t i = 0; i < p->n; i++) d[i] = s[bit_rev_twice(nb, bb, i)]; // src/unordered.rs:967-969
    return CFFT_OK;
}
cfft_status cfft_unordered_from_standard_host(const cfft_plan *p, const void *src, uint64_t count, void *dst)
{
    if (!p || p->kind == KIND_F128 || !src || !dst) return fail(CFFT_EINVAL, "bad argument");
    const unsigned nb = ilog2(p->n), bb = p->kind == KIND_ORDERED ? nb : ilog2(p->base_n);
    const cplx *s = static_cast<const cplx *>(src);
    cplx *d = static_cast<cplx *>(dst);
    for (uint64_t i = 0; i < count && i < p->n; i++) d[bit_rev_twice(nb, bb, i)] = s[i]; // :1019-1022
    if (count != p->n) return fail(CFFT_ELENGTH, "invalid length (src/unordered.rs:1027-1028)");
    return CFFT_OK;
}

} // extern "C"

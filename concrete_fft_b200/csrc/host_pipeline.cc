// host_pipeline.cc -- the *_host entry points: host memory in, host memory out.
//
// This is the literal stand-in for the reference's Plan::fwd(&mut [c64], stack) /
// Plan::inv / fft128 Plan::fwd / inv (one synchronous call on host slices), extended with a
// batch count.  The batch is cut into chunks that flow through three slots, each with its own
// stream and device buffer:  H2D(chunk i+1)  ||  kernels(chunk i)  ||  D2H(chunk i-1).
// Pinned (or cudaHostRegister'ed) caller memory is DMA'd directly; pageable memory is staged
// through internal pinned buffers.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/cfft_b200.h"
#include "plan.h"

using namespace cfft;

namespace {

constexpr int kSlots = 4; // capacity; the number in use comes from pipe_slots()
size_t pipe_chunk_bytes()
{
    static const size_t v = [] {
        const char *e = getenv("CFFT_B200_PIPE_CHUNK_MB");
        const long mb = e ? atol(e) : 32;
        return size_t(mb > 0 ? mb : 32) << 20;
    }();
    return v;
}
int pipe_slots()
{
    static const int v = [] {
        const char *e = getenv("CFFT_B200_PIPE_SLOTS");
        const int k = e ? atoi(e) : 3;
        return k < 1 ? 1 : (k > kSlots ? kSlots : k);
    }();
    return v;
}

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void *dev = nullptr;
    void *pinned = nullptr;
    size_t dev_bytes = 0, pinned_bytes = 0;
    // pending copy-out from the pinned staging buffer (pageable callers only)
    bool pending = false;
    size_t pend_row0 = 0, pend_rows = 0;
};

struct PipeCtx {
    int device = -1;
    Slot slot[kSlots];
};

std::mutex g_pool_mu;
std::vector<PipeCtx *> g_pool;

PipeCtx *acquire_ctx(int device)
{
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t i = 0; i < g_pool.size(); i++)
            if (g_pool[i]->device == device) {
                PipeCtx *c = g_pool[i];
                g_pool.erase(g_pool.begin() + long(i));
                return c;
            }
    }
    PipeCtx *c = new PipeCtx;
    c->device = device;
    return c;
}
void release_ctx(PipeCtx *c)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    g_pool.push_back(c);
}

cudaError_t ensure_slot(Slot &s, size_t dev_bytes, size_t pinned_bytes)
{
    cudaError_t e;
    if (!s.stream) {
        if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    if (s.dev_bytes < dev_bytes) {
        if (s.dev) cudaFree(s.dev);
        s.dev = nullptr;
        s.dev_bytes = 0;
        if ((e = cudaMalloc(&s.dev, dev_bytes)) != cudaSuccess) return e;
        s.dev_bytes = dev_bytes;
    }
    if (s.pinned_bytes < pinned_bytes) {
        if (s.pinned) cudaFreeHost(s.pinned);
        s.pinned = nullptr;
        s.pinned_bytes = 0;
        if ((e = cudaHostAlloc(&s.pinned, pinned_bytes, cudaHostAllocDefault)) != cudaSuccess) return e;
        s.pinned_bytes = pinned_bytes;
    }
    return cudaSuccess;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

thread_local std::string g_err;

// planes: host base pointers; each plane holds batch rows of row_bytes.
// op: 0 fwd, 1 inv, 2 fwd then inv
cfft_status run_pipeline(const cfft_plan *plan, void *const *planes, int nplanes, size_t row_bytes, uint64_t batch,
                         int op, std::string &err)
{
    if (batch == 0) return CFFT_OK;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    if (prev_dev != plan->device && cudaSetDevice(plan->device) != cudaSuccess) {
        err = "cudaSetDevice failed";
        return CFFT_ECUDA;
    }
    bool pinned = true;
    for (int i = 0; i < nplanes; i++) pinned = pinned && is_pinned(planes[i]);
    // c64 kernels use 128-bit accesses: a caller slice that is only 8-byte aligned (Rust's Complex64
    // alignment) cannot be touched in place by the zero-copy path; DMA into our aligned buffers is fine
    bool aligned16 = true;
    for (int i = 0; i < nplanes; i++) aligned16 = aligned16 && (reinterpret_cast<uintptr_t>(planes[i]) & 15) == 0;

    const int nslots = pipe_slots();
    size_t rows_per_chunk = pipe_chunk_bytes() / (row_bytes * size_t(nplanes));
    if (rows_per_chunk < 1) rows_per_chunk = 1;
    if (rows_per_chunk > batch) rows_per_chunk = size_t(batch);
    const size_t chunk_plane_bytes = rows_per_chunk * row_bytes;
    const size_t chunk_bytes = chunk_plane_bytes * size_t(nplanes);

    PipeCtx *ctx = acquire_ctx(plan->device);
    cudaError_t e = cudaSuccess;
    const char *what = "";

    // Small calls (the literal one-polynomial Plan::fwd): skip the copy engines.  The kernels read and
    // write the caller's pinned memory -- or a pinned bounce buffer for pageable memory -- directly
    // over PCIe (zero copy), which saves two DMA submissions and their latencies per call.
    const size_t total_bytes = size_t(batch) * row_bytes * size_t(nplanes);
    static const size_t zero_copy_max = [] {
        const char *ev = getenv("CFFT_B200_ZERO_COPY_MAX_KB");
        return size_t(ev ? atol(ev) : 512) << 10;
    }();
    if (total_bytes <= zero_copy_max) {
        Slot &s = ctx->slot[0];
        what = "zero-copy setup";
        const bool in_place = pinned && (aligned16 || plan->kind == KIND_F128);
        e = ensure_slot(s, 0, in_place ? 0 : total_bytes);
        char *base[4] = {nullptr, nullptr, nullptr, nullptr};
        const size_t plane_bytes = size_t(batch) * row_bytes;
        if (e == cudaSuccess) {
            for (int pl = 0; pl < nplanes; pl++) {
                if (in_place) base[pl] = static_cast<char *>(planes[pl]);
                else {
                    base[pl] = static_cast<char *>(s.pinned) + size_t(pl) * plane_bytes;
                    std::memcpy(base[pl], planes[pl], plane_bytes);
                }
            }
            what = "zero-copy launch";
            for (int pass = 0; pass < (op == 2 ? 2 : 1) && e == cudaSuccess; pass++) {
                const bool inverse = (op == 1) || (op == 2 && pass == 1);
                if (plan->kind == KIND_F128)
                    e = launch_f128(plan, inverse, reinterpret_cast<double *>(base[0]), reinterpret_cast<double *>(base[1]),
                                    reinterpret_cast<double *>(base[2]), reinterpret_cast<double *>(base[3]), batch, s.stream);
                else
                    e = launch_c64(plan, inverse, reinterpret_cast<double2 *>(base[0]), batch, s.stream);
            }
        }
        if (e == cudaSuccess) {
            what = "zero-copy sync";
            e = cudaStreamSynchronize(s.stream);
        }
        if (e == cudaSuccess && !in_place)
            for (int pl = 0; pl < nplanes; pl++) std::memcpy(planes[pl], base[pl], plane_bytes);
        release_ctx(ctx);
        if (prev_dev >= 0 && prev_dev != plan->device) cudaSetDevice(prev_dev);
        if (e != cudaSuccess) {
            err = std::string(what) + ": " + cudaGetErrorString(e);
            return CFFT_ECUDA;
        }
        return CFFT_OK;
    }
    auto flush_pending = [&](Slot &s) {
        if (!s.pending) return;
        for (int pl = 0; pl < nplanes; pl++)
            std::memcpy(static_cast<char *>(planes[pl]) + s.pend_row0 * row_bytes,
                        static_cast<char *>(s.pinned) + size_t(pl) * chunk_plane_bytes, s.pend_rows * row_bytes);
        s.pending = false;
    };

    const size_t nchunks = (size_t(batch) + rows_per_chunk - 1) / rows_per_chunk;
    for (size_t c = 0; c < nchunks && e == cudaSuccess; c++) {
        Slot &s = ctx->slot[c % size_t(nslots)];
        what = "pipeline slot setup";
        if ((e = ensure_slot(s, chunk_bytes, pinned ? 0 : chunk_bytes)) != cudaSuccess) break;
        const size_t row0 = c * rows_per_chunk;
        const size_t rows = (row0 + rows_per_chunk <= batch) ? rows_per_chunk : size_t(batch) - row0;
        if (!pinned) {
            // the slot's previous D2H must have landed before its staging buffer is reused
            what = "pipeline event sync";
            if (c >= size_t(nslots) && (e = cudaEventSynchronize(s.done)) != cudaSuccess) break;
            flush_pending(s);
            for (int pl = 0; pl < nplanes; pl++)
                std::memcpy(static_cast<char *>(s.pinned) + size_t(pl) * chunk_plane_bytes,
                            static_cast<char *>(planes[pl]) + row0 * row_bytes, rows * row_bytes);
        }
        what = "pipeline H2D";
        for (int pl = 0; pl < nplanes && e == cudaSuccess; pl++) {
            const void *src = pinned ? static_cast<char *>(planes[pl]) + row0 * row_bytes
                                     : static_cast<char *>(s.pinned) + size_t(pl) * chunk_plane_bytes;
            e = cudaMemcpyAsync(static_cast<char *>(s.dev) + size_t(pl) * chunk_plane_bytes, src, rows * row_bytes,
                                cudaMemcpyHostToDevice, s.stream);
        }
        if (e != cudaSuccess) break;
        what = "pipeline kernel launch";
        char *d = static_cast<char *>(s.dev);
        for (int pass = 0; pass < (op == 2 ? 2 : 1) && e == cudaSuccess; pass++) {
            const bool inverse = (op == 1) || (op == 2 && pass == 1);
            if (plan->kind == KIND_F128)
                e = launch_f128(plan, inverse, reinterpret_cast<double *>(d),
                                reinterpret_cast<double *>(d + chunk_plane_bytes),
                                reinterpret_cast<double *>(d + 2 * chunk_plane_bytes),
                                reinterpret_cast<double *>(d + 3 * chunk_plane_bytes), rows, s.stream);
            else
                e = launch_c64(plan, inverse, reinterpret_cast<double2 *>(d), rows, s.stream);
        }
        if (e != cudaSuccess) break;
        what = "pipeline D2H";
        for (int pl = 0; pl < nplanes && e == cudaSuccess; pl++) {
            void *dst = pinned ? static_cast<char *>(planes[pl]) + row0 * row_bytes
                               : static_cast<char *>(s.pinned) + size_t(pl) * chunk_plane_bytes;
            e = cudaMemcpyAsync(dst, d + size_t(pl) * chunk_plane_bytes, rows * row_bytes, cudaMemcpyDeviceToHost,
                                s.stream);
        }
        if (e != cudaSuccess) break;
        if (!pinned) {
            s.pending = true;
            s.pend_row0 = row0;
            s.pend_rows = rows;
            e = cudaEventRecord(s.done, s.stream); // staging-buffer reuse waits on this
        }
    }
    // drain the slots this call used
    const int used = int(nchunks < size_t(nslots) ? nchunks : size_t(nslots));
    for (int i = 0; i < used; i++) {
        Slot &s = ctx->slot[i];
        if (!s.stream) continue;
        cudaError_t e2 = cudaStreamSynchronize(s.stream);
        if (e == cudaSuccess && e2 != cudaSuccess) {
            e = e2;
            what = "pipeline drain";
        }
        if (e == cudaSuccess) flush_pending(s);
        s.pending = false;
    }
    release_ctx(ctx);
    if (prev_dev >= 0 && prev_dev != plan->device) cudaSetDevice(prev_dev);
    if (e != cudaSuccess) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return CFFT_ECUDA;
    }
    return CFFT_OK;
}

// out_host[r] (+)= the negacyclic product step of launch_c64_poly_mul for host-resident polynomials: per chunk, the k terms of
// its rows go up (k * 16 n bytes per row), one fused kernel runs against the device-resident Fourier-domain operand b, and
// the result polynomials come back (16 n bytes per row) -- k + 1 row-sizes over PCIe for k transforms forward and one back,
// where the plain host entry points move 2 per transform.  Same three-slot overlap as run_pipeline.
cfft_status run_poly_pipeline(const cfft_plan *plan, const long long *a_host, uint64_t kterms, const double2 *b_dev, uint64_t b_row_stride,
                              long long *out_host, uint64_t batch, uint32_t flags, std::string &err)
{
    if (batch == 0) return CFFT_OK;
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    if (prev_dev != plan->device && cudaSetDevice(plan->device) != cudaSuccess) {
        err = "cudaSetDevice failed";
        return CFFT_ECUDA;
    }
    const size_t in_row = size_t(kterms) * 2 * plan->n * sizeof(long long), out_row = size_t(2) * plan->n * sizeof(long long);
    const bool pinned = is_pinned(a_host) && is_pinned(out_host);
    const bool accumulate = (flags & 2u) != 0; // the device needs the old output polynomials too
    const int nslots = pipe_slots();
    size_t rows_per_chunk = pipe_chunk_bytes() / (in_row + out_row);
    if (rows_per_chunk < 1) rows_per_chunk = 1;
    if (rows_per_chunk > batch) rows_per_chunk = size_t(batch);
    const size_t in_bytes = rows_per_chunk * in_row, out_bytes = rows_per_chunk * out_row;

    PipeCtx *ctx = acquire_ctx(plan->device);
    cudaError_t e = cudaSuccess;
    const char *what = "";
    auto flush_pending = [&](Slot &s) {
        if (!s.pending) return;
        std::memcpy(reinterpret_cast<char *>(out_host) + s.pend_row0 * out_row, static_cast<char *>(s.pinned) + in_bytes, s.pend_rows * out_row);
        s.pending = false;
    };
    const size_t nchunks = (size_t(batch) + rows_per_chunk - 1) / rows_per_chunk;
    for (size_t c = 0; c < nchunks && e == cudaSuccess; c++) {
        Slot &s = ctx->slot[c % size_t(nslots)];
        what = "poly pipeline slot setup";
        if ((e = ensure_slot(s, in_bytes + out_bytes, pinned ? 0 : in_bytes + out_bytes)) != cudaSuccess) break;
        const size_t row0 = c * rows_per_chunk;
        const size_t rows = (row0 + rows_per_chunk <= batch) ? rows_per_chunk : size_t(batch) - row0;
        const char *src_a = reinterpret_cast<const char *>(a_host) + row0 * in_row;
        char *dst_o = reinterpret_cast<char *>(out_host) + row0 * out_row;
        char *d_in = static_cast<char *>(s.dev), *d_out = d_in + in_bytes;
        if (!pinned) {
            what = "poly pipeline event sync";
            if (c >= size_t(nslots) && (e = cudaEventSynchronize(s.done)) != cudaSuccess) break;
            flush_pending(s);
            std::memcpy(s.pinned, src_a, rows * in_row);
            if (accumulate) std::memcpy(static_cast<char *>(s.pinned) + in_bytes, dst_o, rows * out_row);
        }
        what = "poly pipeline H2D";
        e = cudaMemcpyAsync(d_in, pinned ? src_a : static_cast<const char *>(s.pinned), rows * in_row, cudaMemcpyHostToDevice, s.stream);
        if (e == cudaSuccess && accumulate)
            e = cudaMemcpyAsync(d_out, pinned ? dst_o : static_cast<char *>(s.pinned) + in_bytes, rows * out_row, cudaMemcpyHostToDevice, s.stream);
        if (e != cudaSuccess) break;
        what = "poly pipeline kernel launch";
        e = launch_c64_poly_mul(plan, reinterpret_cast<const long long *>(d_in), kterms, b_dev + row0 * b_row_stride, b_row_stride,
                                reinterpret_cast<long long *>(d_out), rows, flags, s.stream);
        if (e != cudaSuccess) break;
        what = "poly pipeline D2H";
        e = cudaMemcpyAsync(pinned ? dst_o : static_cast<char *>(s.pinned) + in_bytes, d_out, rows * out_row, cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) break;
        if (!pinned) {
            s.pending = true;
            s.pend_row0 = row0;
            s.pend_rows = rows;
            e = cudaEventRecord(s.done, s.stream);
        }
    }
    const int used = int(nchunks < size_t(nslots) ? nchunks : size_t(nslots));
    for (int i = 0; i < used; i++) {
        Slot &s = ctx->slot[i];
        if (!s.stream) continue;
        cudaError_t e2 = cudaStreamSynchronize(s.stream);
        if (e == cudaSuccess && e2 != cudaSuccess) {
            e = e2;
            what = "poly pipeline drain";
        }
        if (e == cudaSuccess) flush_pending(s);
        s.pending = false;
    }
    release_ctx(ctx);
    if (prev_dev >= 0 && prev_dev != plan->device) cudaSetDevice(prev_dev);
    if (e != cudaSuccess) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return CFFT_ECUDA;
    }
    return CFFT_OK;
}

} // namespace

// defined in api.cc
extern "C" const char *cfft_last_error(void);
namespace cfft { cfft_status set_last_error(cfft_status st, const std::string &msg); }

extern "C" {

static cfft_status c64_host(const cfft_plan *p, void *host_buf, uint64_t len, uint64_t batch, int op)
{
    if (!p || p->kind == KIND_F128) return set_last_error(CFFT_EINVAL, "not a c64 plan");
    if (len != batch * p->n) return set_last_error(CFFT_ELENGTH, "buffer length != batch * fft size (src/unordered.rs:827)");
    if (!host_buf && len) return set_last_error(CFFT_EINVAL, "null buffer");
    std::string err;
    void *planes[1] = {host_buf};
    cfft_status st = run_pipeline(p, planes, 1, size_t(p->n) * sizeof(cplx), batch, op, err);
    return st == CFFT_OK ? st : set_last_error(st, err);
}

cfft_status cfft_c64_fwd_host(const cfft_plan *p, void *host_buf, uint64_t len, uint64_t batch)
{
    return c64_host(p, host_buf, len, batch, 0);
}
cfft_status cfft_c64_inv_host(const cfft_plan *p, void *host_buf, uint64_t len, uint64_t batch)
{
    return c64_host(p, host_buf, len, batch, 1);
}
cfft_status cfft_c64_fwd_inv_host(const cfft_plan *p, void *host_buf, uint64_t len, uint64_t batch)
{
    return c64_host(p, host_buf, len, batch, 2);
}

static cfft_status f128_host(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t len,
                             uint64_t batch, int op)
{
    if (!p || p->kind != KIND_F128) return set_last_error(CFFT_EINVAL, "not an fft128 plan");
    if (len != batch * p->n) return set_last_error(CFFT_ELENGTH, "buffer length != batch * fft size (src/fft128/mod.rs:1912-1915)");
    if (len && (!re0 || !re1 || !im0 || !im1)) return set_last_error(CFFT_EINVAL, "null buffer");
    std::string err;
    void *planes[4] = {re0, re1, im0, im1};
    cfft_status st = run_pipeline(p, planes, 4, size_t(p->n) * sizeof(double), batch, op, err);
    return st == CFFT_OK ? st : set_last_error(st, err);
}

cfft_status cfft_f128_fwd_host(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t len,
                               uint64_t batch)
{
    return f128_host(p, re0, re1, im0, im1, len, batch, 0);
}
cfft_status cfft_f128_inv_host(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t len,
                               uint64_t batch)
{
    return f128_host(p, re0, re1, im0, im1, len, batch, 1);
}

cfft_status cfft_f128_fwd_inv_host(const cfft_plan *p, double *re0, double *re1, double *im0, double *im1, uint64_t len,
                                   uint64_t batch)
{
    return f128_host(p, re0, re1, im0, im1, len, batch, 2);
}

// ---- one host call, several GPUs (north_star item 5 inside the library) ---------------------------------------------
// `plans` are replicas of ONE plan on the devices that should share the work (cfft_plan_clone_to_device; two replicas may
// sit on the same device).  The batch is cut into nplans contiguous row ranges and every range runs through its replica's
// own chunked H2D / kernels / D2H pipeline on a host thread of its own: no collective, no peer traffic -- polynomials are
// independent (Plan::fwd is one polynomial per call).  What limits it is the host side of PCIe, not the GPUs.
static bool same_transform(const cfft_plan *a, const cfft_plan *b)
{
    return a->kind == b->kind && a->n == b->n && (a->kind == KIND_F128 || (a->algo == b->algo && a->base_n == b->base_n && a->allow_large == b->allow_large));
}

static cfft_status host_multi(const cfft_plan *const *plans, int nplans, void *const *planes, int nplanes, size_t row_bytes, uint64_t batch, int op)
{
    if (nplans == 1) {
        std::string err;
        cfft_status st = run_pipeline(plans[0], planes, nplanes, row_bytes, batch, op, err);
        return st == CFFT_OK ? st : set_last_error(st, err);
    }
    std::vector<std::thread> workers;
    std::vector<cfft_status> status(size_t(nplans), CFFT_OK);
    std::vector<std::string> errs{size_t(nplans)};
    const uint64_t base = batch / uint64_t(nplans), rem = batch % uint64_t(nplans);
    uint64_t row0 = 0;
    for (int i = 0; i < nplans; i++) {
        const uint64_t rows = base + (uint64_t(i) < rem ? 1 : 0);
        if (rows == 0) continue;
        workers.emplace_back([=, &status, &errs] {
            void *shifted[4] = {nullptr, nullptr, nullptr, nullptr};
            for (int pl = 0; pl < nplanes; pl++) shifted[pl] = static_cast<char *>(planes[pl]) + row0 * row_bytes;
            status[size_t(i)] = run_pipeline(plans[i], shifted, nplanes, row_bytes, rows, op, errs[size_t(i)]);
        });
        row0 += rows;
    }
    for (std::thread &w : workers) w.join();
    for (int i = 0; i < nplans; i++)
        if (status[size_t(i)] != CFFT_OK) return set_last_error(status[size_t(i)], "replica " + std::to_string(i) + " (cuda:" + std::to_string(plans[i]->device) + "): " + errs[size_t(i)]);
    return CFFT_OK;
}

static cfft_status check_replicas(const cfft_plan *const *plans, int nplans, bool f128)
{
    if (!plans || nplans < 1 || nplans > 64) return set_last_error(CFFT_EINVAL, "need 1 .. 64 plan replicas");
    for (int i = 0; i < nplans; i++) {
        if (!plans[i] || (plans[i]->kind == KIND_F128) != f128) return set_last_error(CFFT_EINVAL, f128 ? "not an fft128 plan" : "not a c64 plan");
        if (!same_transform(plans[0], plans[i])) return set_last_error(CFFT_EINVAL, "replicas must describe the same transform (kind, n, base algorithm, base size)");
    }
    return CFFT_OK;
}

cfft_status cfft_c64_host_multi(const cfft_plan *const *plans, int nplans, int op, void *host_buf, uint64_t len, uint64_t batch)
{
    cfft_status st = check_replicas(plans, nplans, false);
    if (st != CFFT_OK) return st;
    if (op < 0 || op > 2) return set_last_error(CFFT_EINVAL, "op: 0 fwd, 1 inv, 2 fwd then inv");
    if (len != batch * plans[0]->n) return set_last_error(CFFT_ELENGTH, "buffer length != batch * fft size (src/unordered.rs:827)");
    if (!host_buf && len) return set_last_error(CFFT_EINVAL, "null buffer");
    void *planes[1] = {host_buf};
    return host_multi(plans, nplans, planes, 1, size_t(plans[0]->n) * sizeof(cplx), batch, op);
}

cfft_status cfft_f128_host_multi(const cfft_plan *const *plans, int nplans, int op, double *re0, double *re1, double *im0, double *im1,
                                 uint64_t len, uint64_t batch)
{
    cfft_status st = check_replicas(plans, nplans, true);
    if (st != CFFT_OK) return st;
    if (op < 0 || op > 2) return set_last_error(CFFT_EINVAL, "op: 0 fwd, 1 inv, 2 fwd then inv");
    if (len != batch * plans[0]->n) return set_last_error(CFFT_ELENGTH, "buffer length != batch * fft size (src/fft128/mod.rs:1912-1915)");
    if (len && (!re0 || !re1 || !im0 || !im1)) return set_last_error(CFFT_EINVAL, "null buffer");
    void *planes[4] = {re0, re1, im0, im1};
    return host_multi(plans, nplans, planes, 4, size_t(plans[0]->n) * sizeof(double), batch, op);
}

cfft_status cfft_c64_poly_mul_host(const cfft_plan *p, const int64_t *a_host, uint64_t k_terms, const void *b_dev, uint64_t b_row_stride,
                                   int64_t *out_host, uint64_t batch, uint32_t flags)
{
    if (!p || p->kind == KIND_F128) return set_last_error(CFFT_EINVAL, "not a c64 plan");
    if (k_terms == 0) return set_last_error(CFFT_EINVAL, "k_terms must be >= 1");
    if (flags & ~3u) return set_last_error(CFFT_EINVAL, "unknown flag bits");
    if (batch && (!a_host || !b_dev || !out_host)) return set_last_error(CFFT_EINVAL, "null buffer");
    if (reinterpret_cast<uintptr_t>(b_dev) & 15) return set_last_error(CFFT_EINVAL, "b must be 16-byte aligned device memory");
    if ((reinterpret_cast<uintptr_t>(a_host) | reinterpret_cast<uintptr_t>(out_host)) & 7) return set_last_error(CFFT_EINVAL, "polynomials must be 8-byte aligned");
    if (b_row_stride != 0 && b_row_stride < k_terms * p->n) return set_last_error(CFFT_EINVAL, "b_row_stride must be 0 (b shared by every row) or >= k_terms * n");
    std::string err;
    cfft_status st = run_poly_pipeline(p, reinterpret_cast<const long long *>(a_host), k_terms, static_cast<const double2 *>(b_dev), b_row_stride,
                                       reinterpret_cast<long long *>(out_host), batch, flags, err);
    return st == CFFT_OK ? st : set_last_error(st, err);
}

cfft_status cfft_unordered_fwd_monomial_host(const cfft_plan *p, uint64_t degree, void *host_buf, uint64_t len)
{
    if (!p || p->kind != KIND_UNORDERED) return set_last_error(CFFT_EINVAL, "not an unordered plan");
    if (len != p->n) return set_last_error(CFFT_ELENGTH, "buffer length != fft size (src/unordered.rs:858)");
    if (degree >= p->n) return set_last_error(CFFT_EINVAL, "degree must be < n (src/unordered.rs:859)");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) cudaSetDevice(p->device);
    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, p->n * sizeof(cplx));
    if (e == cudaSuccess) e = launch_monomial(p, degree, static_cast<double2 *>(d), nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(host_buf, d, p->n * sizeof(cplx), cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (prev >= 0 && prev != p->device) cudaSetDevice(prev);
    if (e != cudaSuccess) return set_last_error(CFFT_ECUDA, std::string("fwd_monomial: ") + cudaGetErrorString(e));
    return CFFT_OK;
}

} // extern "C"

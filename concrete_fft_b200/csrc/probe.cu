// probe.cu -- measured FP64 issue rate of the device: the denominator of the fft128 roofline.
//
// The fft128 butterflies (src/fft128/mod.rs:310-346 over f128_ops.rs:302-356, 837-841) are 94 FP64
// instructions each -- 78 DADD, 12 DFMA, 4 DMUL -- so what bounds the kernel is how many FP64
// instructions per second the SMs can issue, not FMA flops.  bench.py calls this probe on the box it
// runs on and reports fft128's achieved instruction rate against the number it returns, instead of
// the nominal 64 lanes x SM count x max clock.
#include <cuda_runtime.h>

#include "../../include/cfft_b200.h"

namespace {

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct ProbeClock {
    unsigned long long c0, c1, g0, g1;
};

// 8 independent dependency chains per thread, 16 warps per SM: latency (8 cycles at most) is covered many times over
template <int MIX>
__global__ void __launch_bounds__(256) fp64_issue_kernel(double *sink, int iters, ProbeClock *clk)
{
    double a[8];
    const double b = 1.0 + 1e-9 * threadIdx.x, c = 1e-12;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = 1.0 + j + threadIdx.x * 1e-3;
    unsigned long long c0 = 0, g0 = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        c0 = clock64();
        g0 = global_ns();
    }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MIX == 0) a[j] = __fma_rn(a[j], b, c);
            else if (MIX == 1) a[j] = __dadd_rn(a[j], c);
            else a[j] = (j == 0) ? __fma_rn(a[j], b, c) : __dadd_rn(a[j], c); // 1 : 7, close to fft128's 16 : 78
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        clk->c0 = c0;
        clk->g0 = g0;
        clk->c1 = clock64();
        clk->g1 = global_ns();
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    if (s == 123.456) sink[0] = s; // never true: keeps the chains alive
}

} // namespace

extern "C" cfft_status cfft_probe_fp64_issue_rate(int device, double *dfma_per_s, double *dadd_per_s, double *mix_per_s,
                                                  double *sm_mhz, int *sm_count)
{
    int prev = -1;
    if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return CFFT_ECUDA;
    cfft_status st = CFFT_OK;
    int sms = 0;
    double *sink = nullptr;
    ProbeClock *clk = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    double rate[3] = {0, 0, 0}, mhz = 0;
    auto ok = [&](cudaError_t e) {
        if (e != cudaSuccess) st = CFFT_ECUDA;
        return e == cudaSuccess;
    };
    if (ok(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device)) && ok(cudaMalloc(&sink, 64)) &&
        ok(cudaMalloc(&clk, sizeof(ProbeClock))) && ok(cudaEventCreate(&e0)) && ok(cudaEventCreate(&e1))) {
        const int grid = sms * 2, iters = 1 << 15; // 2 x 256 threads per SM = 4 warps per scheduler
        for (int mix = 0; mix < 3 && st == CFFT_OK; mix++) {
            float best = 1e30f;
            for (int rep = 0; rep < 4 && st == CFFT_OK; rep++) { // first repetition warms the clocks up
                ok(cudaEventRecord(e0));
                if (mix == 0) fp64_issue_kernel<0><<<grid, 256>>>(sink, iters, clk);
                else if (mix == 1) fp64_issue_kernel<1><<<grid, 256>>>(sink, iters, clk);
                else fp64_issue_kernel<2><<<grid, 256>>>(sink, iters, clk);
                ok(cudaEventRecord(e1));
                ok(cudaEventSynchronize(e1));
                float ms = 0;
                if (ok(cudaEventElapsedTime(&ms, e0, e1)) && rep > 0 && ms < best) best = ms;
            }
            rate[mix] = double(grid) * 256.0 * 8.0 * iters / (double(best) * 1e-3);
            ProbeClock h;
            if (ok(cudaMemcpy(&h, clk, sizeof(h), cudaMemcpyDeviceToHost)) && h.g1 > h.g0) mhz = double(h.c1 - h.c0) / double(h.g1 - h.g0) * 1e3;
        }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (sink) cudaFree(sink);
    if (clk) cudaFree(clk);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (st != CFFT_OK) return st;
    if (dfma_per_s) *dfma_per_s = rate[0];
    if (dadd_per_s) *dadd_per_s = rate[1];
    if (mix_per_s) *mix_per_s = rate[2];
    if (sm_mhz) *sm_mhz = mhz;
    if (sm_count) *sm_count = sms;
    return CFFT_OK;
}

// ubench.cu -- B200 micro-benchmarks behind the design decisions of DESIGN.md (run on the GPU box, results under profiles/):
//   1. FP64 issue rate (DFMA / DADD / DMUL chains)        -> the measured denominator of the fft128 roofline
//   2. warp shuffle vs 128-bit shared-memory store + load -> is a shuffle transpose cheaper than the XOR-swizzled tile?
//   3. L2 bandwidth: read / write / copy of L2-resident buffers with L1 bypassed -> the cap of any two-pass (n >= 2^14) schedule
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct Clk {
    unsigned long long c0, c1, g0, g1;
};

template <int OP> __global__ void __launch_bounds__(256) fp64_rate(double *out, int iters, Clk *clk)
{
    double a[8];
    const double b = 1.0 + 1e-9 * threadIdx.x, c = 1e-12;
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = 1.0 + j + threadIdx.x * 1e-3;
    unsigned long long c0 = 0, g0 = 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) { c0 = clock64(); g0 = gtime(); }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (OP == 0) a[j] = __fma_rn(a[j], b, c);
            else if (OP == 1) a[j] = __dadd_rn(a[j], c);
            else if (OP == 2) a[j] = __dmul_rn(a[j], b);
            else a[j] = (j & 1) ? __fma_rn(a[j], b, c) : __dadd_rn(a[j], c); // the fft128 mix is mostly DADD + some DFMA
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { clk->c0 = c0; clk->g0 = g0; clk->c1 = clock64(); clk->g1 = gtime(); }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) shfl_rate(unsigned *out, int iters)
{
    unsigned v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = threadIdx.x * 17u + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __shfl_xor_sync(0xFFFFFFFFu, v[j], 1 + (j & 7)) + 1u;
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) s += v[j];
    if (s == 0xDEADBEEF) out[0] = s;
}

// 16 STS.128 + 16 LDS.128 per iteration, the access pattern of the 16 x 16 half-warp transpose (XOR swizzle: conflict-free)
__global__ void __launch_bounds__(128) smem_rate(double *out, int iters)
{
    __shared__ double2 s[128 * 16];
    const int hw = threadIdx.x >> 4, l = threadIdx.x & 15;
    double2 *blk = s + hw * 256;
    double2 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = make_double2(threadIdx.x + k, k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) blk[16 * l + (k ^ l)] = v[k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = blk[16 * k + (l ^ k)];
        __syncwarp();
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) acc += v[k].x + v[k].y;
    if (acc == 123.456) out[0] = acc;
}

// the same transpose with shuffles only: 4 exchange stages, 8 c64 (32 words) per stage
__global__ void __launch_bounds__(256) shfl_transpose_rate(double *out, int iters)
{
    const int l = threadIdx.x & 15;
    double2 v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = make_double2(threadIdx.x + k, k);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int b = 1; b < 16; b <<= 1) {
            const bool up = (l & b) != 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                if (k & b) continue;
                // lane without bit b keeps v[k], sends v[k | b]; lane with bit b keeps v[k | b], sends v[k]
                double2 send = up ? v[k] : v[k | b];
                double2 recv;
                recv.x = __shfl_xor_sync(0xFFFFFFFFu, send.x, b);
                recv.y = __shfl_xor_sync(0xFFFFFFFFu, send.y, b);
                if (up) v[k] = recv;
                else v[k | b] = recv;
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) acc += v[k].x + v[k].y;
    if (acc == 123.456) out[0] = acc;
}

__device__ __forceinline__ double2 ld_na(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_na(double2 *p, double2 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// MODE 0 read, 1 write, 2 copy a -> b, 3 read-modify-write in place (the access pattern of an in-place transform pass).
// One sweep = every thread touches elements i, i + n/4, i + n/2, i + 3n/4 (grid = n / 4 / 256 blocks), so each sweep covers
// the buffer exactly once; the start index rotates with the sweep so that no load can be hoisted out of the loop.
template <int MODE> __global__ void __launch_bounds__(256) l2_rate(double2 *a, double2 *b, size_t n, int reps, double *out)
{
    double acc = 0;
    const size_t q = n / 4;
    size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (int r = 0; r < reps; r++) {
        if (MODE == 1) {
#pragma unroll
            for (int u = 0; u < 4; u++) st_na(a + i + u * q, make_double2(r, u));
        } else {
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = ld_na(a + i + u * q);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (MODE == 2) st_na(b + i + u * q, v[u]);
                else if (MODE == 3) st_na(a + i + u * q, make_double2(v[u].x + 1.0, v[u].y));
                else acc += v[u].x + v[u].y;
            }
        }
        i += 256 * 37; // another block's slice next sweep
        if (i >= q) i -= q;
    }
    if (acc == 123.456) out[0] = acc;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, L2 %.0f MiB, clockRate %.0f MHz\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0,
           prop.clockRate / 1e3);
    const int sms = prop.multiProcessorCount;
    double *out;
    Clk *clk;
    CK(cudaMalloc(&out, 64));
    CK(cudaMalloc(&clk, sizeof(Clk)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0;

    // 1. FP64 issue rate
    const char *names[4] = {"DFMA", "DADD", "DMUL", "DADD+DFMA 1:1"};
    for (int warps_per_smsp = 1; warps_per_smsp <= 8; warps_per_smsp *= 2) {
        for (int op = 0; op < 4; op++) {
            const int ctas = sms * warps_per_smsp / 2, iters = 1 << 15; // 256 threads = 8 warps = 2 per SMSP
            const int grid = ctas < sms ? sms : ctas;
            const int threads = ctas < sms ? 128 * warps_per_smsp : 256;
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                if (op == 0) fp64_rate<0><<<grid, threads>>>(out, iters, clk);
                if (op == 1) fp64_rate<1><<<grid, threads>>>(out, iters, clk);
                if (op == 2) fp64_rate<2><<<grid, threads>>>(out, iters, clk);
                if (op == 3) fp64_rate<3><<<grid, threads>>>(out, iters, clk);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
            }
            CK(cudaEventElapsedTime(&ms, e0, e1));
            Clk h;
            CK(cudaMemcpy(&h, clk, sizeof(h), cudaMemcpyDeviceToHost));
            const double mhz = double(h.c1 - h.c0) / double(h.g1 - h.g0) * 1e3;
            const double instr = double(grid) * threads * 8.0 * iters;
            printf("fp64 %-14s %d warp(s)/SMSP: %.3f T thread-instr/s, %.1f lanes/clk/SM at %.0f MHz (SM clock measured in-kernel)\n", names[op],
                   warps_per_smsp, instr / ms / 1e9, instr / (ms * 1e-3) / sms / (mhz * 1e6), mhz);
        }
    }

    // 2. shuffle vs shared memory
    {
        const int grid = sms * 4, iters = 1 << 13;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(e0));
            shfl_rate<<<grid, 256>>>(reinterpret_cast<unsigned *>(out), iters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double winstr = double(grid) * 8 * 8.0 * iters;
        printf("SHFL.BFLY: %.2f G warp-instr/s/SM = %.3f per clk at 1.9 GHz (= %.0f B/clk/SM of 32-bit lanes)\n", winstr / ms / 1e6 / sms,
               winstr / (ms * 1e-3) / sms / 1.9e9, winstr / (ms * 1e-3) / sms / 1.9e9 * 128);
        const int it2 = 1 << 11;
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(e0));
            smem_rate<<<grid * 2, 128>>>(out, it2);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tr = double(grid) * 16.0 * it2; // 256-point transposes (one per half-warp per iteration)
        printf("16x16 c64 transpose via XOR-swizzled smem (16 STS.128 + 16 LDS.128 per thread): %.2f G transposes/s, %.1f clk per transpose per SM, %.0f B/clk/SM (store + load bytes)\n",
               tr / ms / 1e6, 1.9e9 * sms / (tr / (ms * 1e-3)), tr * 8192.0 / (ms * 1e-3) / sms / 1.9e9);
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaEventRecord(e0));
            shfl_transpose_rate<<<grid, 256>>>(out, it2);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("16x16 c64 transpose via 4 stages of SHFL.BFLY (128 shuffles per thread):              %.2f G transposes/s, %.1f clk per transpose per SM\n",
               tr / ms / 1e6, 1.9e9 * sms / (tr / (ms * 1e-3)));
    }

    // 3. L2 / HBM bandwidth, L1 bypassed
    {
        const size_t max_bytes = size_t(2) << 30;
        double2 *a, *b;
        CK(cudaMalloc(&a, max_bytes));
        CK(cudaMalloc(&b, max_bytes));
        CK(cudaMemset(a, 0, max_bytes));
        CK(cudaMemset(b, 0, max_bytes));
        const size_t sizes_mb[] = {8, 16, 24, 32, 48, 64, 96, 128, 2048};
        for (size_t mb : sizes_mb) {
            const size_t n = (mb << 20) / 16;
            const int reps = int((size_t(8) << 30) / (mb << 20));
            for (int mode = 0; mode < 4; mode++) {
                for (int w = 0; w < 2; w++) {
                    CK(cudaEventRecord(e0));
                    if (mode == 0) l2_rate<0><<<unsigned(n / 4 / 256), 256>>>(a, b, n, w ? reps : 1, out);
                    if (mode == 1) l2_rate<1><<<unsigned(n / 4 / 256), 256>>>(a, b, n, w ? reps : 1, out);
                    if (mode == 2) l2_rate<2><<<unsigned(n / 4 / 256), 256>>>(a, b, n, w ? reps : 1, out);
                    if (mode == 3) l2_rate<3><<<unsigned(n / 4 / 256), 256>>>(a, b, n, w ? reps : 1, out);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                }
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = double(mb << 20) * reps * (mode >= 2 ? 2 : 1);
                printf("%-5s buffer %4zu MiB%s x %4d sweeps: %6.0f GB/s (%s bytes)\n", mode == 0 ? "read" : (mode == 1 ? "write" : (mode == 2 ? "copy" : "rmw")), mb,
                       mode == 2 ? " x 2" : "    ", reps, bytes / ms / 1e6, mode >= 2 ? "read + write" : "moved");
            }
        }
    }
    return 0;
}

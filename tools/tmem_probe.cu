// tmem_probe.cu -- tensor memory as a per-thread parking space (design input for the n = 8192 kernel of DESIGN.md 9b):
// every thread of a 256-thread CTA parks 128 32-bit columns (512 B) in its own TMEM lane with tcgen05.st (32x32b shape),
// reads them back with tcgen05.ld, checks the values, and the round trip is timed per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
// NOT yet run on hardware (written at the end of round 1 when the GPU budget was spent); ptxas accepts it.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}

constexpr uint32_t kCols = 256; // 256 columns x 128 lanes x 4 B = 128 KiB: one n = 8192 c64 transform

__global__ void __launch_bounds__(256, 2) probe(int iters, unsigned long long *cycles, unsigned *mismatches)
{
    __shared__ uint32_t tmem_base_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t slot = static_cast<uint32_t>(__cvta_generic_to_shared(&tmem_base_slot));
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(kCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_base_slot;
    // warp w owns TMEM lanes 32 (w % 4) .. + 31; warps w and w + 4 share a quadrant and take columns [0,128) / [128,256)
    const uint32_t taddr = base + ((uint32_t(warp & 3) * 32u) << 16) + uint32_t(warp >> 2) * 128u;
    unsigned bad = 0;
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        uint32_t r[16];
#pragma unroll
        for (int blk = 0; blk < 8; blk++) {
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = threadIdx.x * 4096u + uint32_t(blk * 16 + i) + uint32_t(it) * 7u;
            tmem_st16(taddr + blk * 16, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
        for (int blk = 0; blk < 8; blk++) {
            tmem_ld16(taddr + blk * 16, r);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; i++) bad += r[i] != threadIdx.x * 4096u + uint32_t(blk * 16 + i) + uint32_t(it) * 7u;
        }
    }
    const unsigned long long t1 = clock64();
    if (bad) atomicAdd(mismatches, bad);
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(kCols) : "memory");
}

int main()
{
    unsigned long long *cyc;
    unsigned *bad;
    for (int per_sm = 1; per_sm <= 2; per_sm++) {
        const int ctas = 148 * per_sm, iters = 2000;
        cudaMalloc(&cyc, ctas * sizeof(*cyc));
        cudaMalloc(&bad, sizeof(*bad));
        cudaMemset(bad, 0, sizeof(*bad));
        probe<<<ctas, 256>>>(iters, cyc, bad);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[296];
        unsigned hb = 0;
        cudaMemcpy(h, cyc, ctas * sizeof(*cyc), cudaMemcpyDeviceToHost);
        cudaMemcpy(&hb, bad, sizeof(hb), cudaMemcpyDeviceToHost);
        double mean = 0;
        for (int i = 0; i < ctas; i++) mean += double(h[i]) / ctas;
        const double bytes = 256.0 * 512.0 * iters; // per CTA, each way
        printf("%d CTA(s) per SM: %u mismatches, %.0f cycles per 128 KiB park + fetch round trip, %.1f B/clk/CTA each way\n", per_sm, hb,
               mean / iters, bytes / mean);
        cudaFree(cyc);
        cudaFree(bad);
    }
    return 0;
}

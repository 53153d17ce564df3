// dsmem_probe.cu -- distributed-shared-memory bandwidth between the CTAs of a thread-block cluster
// (design input for a single-pass large-n FFT: the transpose between column and row stages would
// travel over DSMEM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_probe dsmem_probe.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

namespace cg = cooperative_groups;

// every CTA repeatedly writes (or reads) `bytes` of its peers' shared memory with 128-bit accesses
template <bool WRITE>
__global__ void probe(int iters, int bytes, unsigned long long *cycles_out, double *sink)
{
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), csz = cluster.num_blocks();
    double2 *local = reinterpret_cast<double2 *>(smem);
    const int n16 = bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) local[i] = make_double2(i, rank);
    cluster.sync();
    double2 acc = make_double2(0, 0);
    const unsigned long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        // all-to-all: slice j of the buffer goes to / comes from peer (rank + 1 + j) % csz
        for (int i = threadIdx.x; i < n16; i += blockDim.x) {
            const unsigned peer = (rank + 1 + (unsigned(i) * (csz - 1)) / unsigned(n16)) % csz;
            double2 *remote = cluster.map_shared_rank(local, peer);
            if (WRITE) remote[i] = make_double2(it, i);
            else { double2 v = remote[i]; acc.x += v.x; acc.y += v.y; }
        }
        cluster.sync();
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    if (acc.x == 12345.678) sink[0] = acc.x + acc.y;
}

template <bool WRITE> void run(int csz, int threads, int bytes, int iters)
{
    unsigned long long *cyc;
    double *sink;
    const int ctas = 16 * csz; // fewer clusters than SM groups: one CTA per SM
    cudaMalloc(&cyc, ctas * sizeof(*cyc));
    cudaMalloc(&sink, 8);
    cudaFuncSetAttribute(probe<WRITE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (csz > 8) cudaFuncSetAttribute(probe<WRITE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = bytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csz;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaLaunchKernelEx(&cfg, probe<WRITE>, 2, bytes, cyc, sink); // warm-up
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, probe<WRITE>, iters, bytes, cyc, sink);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) {
        printf("cluster %2d %s: launch failed (%s)\n", csz, WRITE ? "write" : "read ", cudaGetErrorString(err));
        return;
    }
    unsigned long long h[256];
    cudaMemcpy(h, cyc, ctas * sizeof(*cyc), cudaMemcpyDeviceToHost);
    unsigned long long mx = 0;
    for (int i = 0; i < ctas; i++) mx = h[i] > mx ? h[i] : mx;
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("cluster %2d threads %4d %s %3d KiB/CTA: %.1f B/clk/SM (max %llu cycles for %d iters), %.1f GB/s per SM by wall time\n",
           csz, threads, WRITE ? "write" : "read ", bytes / 1024, double(bytes) * iters / double(mx), mx, iters,
           double(bytes) * iters / (ms * 1e-3) / 1e9);
    cudaFree(cyc);
    cudaFree(sink);
}

int main()
{
    for (int csz : {2, 4, 8, 16})
        for (int threads : {256, 512}) {
            run<true>(csz, threads, 64 * 1024, 50);
            run<false>(csz, threads, 64 * 1024, 50);
        }
    run<true>(8, 512, 128 * 1024, 50);
    run<false>(8, 512, 128 * 1024, 50);
    return 0;
}

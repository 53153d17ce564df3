// c64_dev.cuh -- device helpers shared by the c64 register kernels (c64_fast.cu, c64_column.cu,
// c64_ord16.cu): streaming 128-bit HBM accesses, twiddle loads, and the 256-point Dif16 base FFT
// of one half-warp (src/dif16.rs:449-827).
#pragma once
#include "c64_math.cuh"

namespace cfft {
namespace dev {

__device__ __forceinline__ c64 ld_stream(const c64 *p)
{
    c64 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(c64 *p, c64 v)
{
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ c64 ld_tw(const c64 *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

template <int R> __device__ __forceinline__ constexpr int brev_c(int k)
{
    return R == 2 ? k : (R == 4 ? ((k & 1) << 1) | (k >> 1) : ((k & 1) << 2) | (k & 2) | (k >> 2));
}

// 256-point base FFT (Dif16: radix-16 s=1 with twiddles, then radix-16 end) of the half-warp
// that owns block `blk`; thread lane16 = p (first pass) = j (second pass).
// FWD selects the butterfly direction only; the table passed in is the direction's table.
// `sw_in` / `sw_out` (0..7) XOR the natural-order shared-memory positions read / written; the standard-order
// kernels use it so that the transposing pass that follows / precedes is bank-conflict free.
template <bool FWD, bool G_IN, bool G_OUT>
__device__ __forceinline__ void base256(const c64 *__restrict__ src, c64 *__restrict__ sm_blk, c64 *__restrict__ dst,
                                        const c64 *__restrict__ tw_planar, int lane16, c64 (&v)[16], int sw_in = 0, int sw_out = 0)
{
    const unsigned hmask = 0xFFFFu << (threadIdx.x & 16); // the 16 lanes that own this block
    // pass 1: x[p + 16k] -> y[16p + k] = w[p + 16k] * DFT16(x)_k       src/dif16.rs:449-623
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = G_IN ? ld_stream(src + lane16 + 16 * k) : src[(lane16 + 16 * k) ^ sw_in];
    bf16<FWD>(v);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(ld_tw(tw_planar + lane16 + 16 * k), v[k]);
    __syncwarp(hmask); // the half-warp has finished reading its block
#pragma unroll
    for (int k = 0; k < 16; k++) sm_blk[16 * lane16 + (k ^ lane16)] = v[k]; // XOR swizzle: conflict-free
    __syncwarp(hmask);
    // pass 2: terminal radix-16 on y[j + 16k]                           src/dif16.rs:649-827
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = sm_blk[16 * k + (lane16 ^ k)];
    bf16<FWD>(v);
    if (G_OUT) {
#pragma unroll
        for (int k = 0; k < 16; k++) st_stream(dst + lane16 + 16 * k, v[k]);
    } else {
        __syncwarp(hmask); // swizzled data consumed by the whole half-warp before natural-order overwrite
#pragma unroll
        for (int k = 0; k < 16; k++) dst[(lane16 + 16 * k) ^ sw_out] = v[k];
    }
}

} // namespace dev
} // namespace cfft

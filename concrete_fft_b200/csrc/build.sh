#!/bin/bash
# Builds concrete_fft_b200/libcfft_b200.so for sm_100a (nvcc cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
OUT=../libcfft_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
HOSTFLAGS="-Xcompiler -fPIC,-ffp-contract=off,-O2,-Wall"
# --fmad=false: no implicit FMA contraction anywhere in device code; every fused op in the
# kernels is an explicit __fma_rn (bit-exactness with the reference depends on it).
CUFLAGS="-std=c++17 -O3 -lineinfo --fmad=false $ARCH $HOSTFLAGS"
mkdir -p _obj
pids=()
for f in c64_tile.cu c64_regs.cu c64_fast.cu c64_poly.cu c64_ord16.cu c64_column.cu c64_colpipe.cu c64_tmem.cu f128.cu f128_ops.cu probe.cu; do
  $NVCC $CUFLAGS ${PTXAS_V:+-Xptxas -v} -c $f -o _obj/${f%.cu}.o &
  pids+=($!)
done
for f in api.cc host_pipeline.cc tables.cc; do
  $NVCC -std=c++17 -O2 $HOSTFLAGS -x cu $ARCH -c $f -o _obj/${f%.cc}.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared $ARCH -o $OUT _obj/*.o -cudart static
echo "built $(realpath $OUT)"

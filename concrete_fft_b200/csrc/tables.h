// tables.h -- host-side twiddle / permutation tables of a plan (computed once, uploaded once).
//
// The values must be the reference's, bit for bit, because the kernels reproduce the
// reference's butterflies exactly; so the generators follow
//   sincospi64                src/fft_simd.rs:237-296
//   init_wt                   src/fft_simd.rs:298-321
//   unordered init_twiddles   src/unordered.rs:349-389   (scalar layout, complex_per_reg = 1)
//   f128 sincospi             src/fft128/f128_ops.rs:514-618
//   init_negacyclic_twiddles  src/fft128/mod.rs:1805-1828
//   bit_rev_twice(_inv)       src/unordered.rs:1039-1059
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace cfft {

struct cplx { double re, im; };

void sincospi64(double a, double &s, double &c);

// 2n entries each; first half planar (w[p + k n/r]), second half interleaved (w[n + r p + k]).
void init_wt(size_t r, size_t n, cplx *w, cplx *w_inv);

// fwd: [top level ... lower levels][base init_wt table (2 base_n)]  -- n + base_n entries
// inv: [base table][lower levels ... top level]                     -- mirrored from the end
void init_unordered_twiddles(size_t n, size_t base_n, size_t base_r, std::vector<cplx> &w,
                             std::vector<cplx> &w_inv);

int algo_radix(int algo);
bool algo_is_dit(int algo);
unsigned ilog2(uint64_t n);
bool is_pow2(uint64_t n);
int top_radix(uint64_t n, uint64_t base_n); // src/unordered.rs:407-413

uint64_t bit_rev(unsigned nbits, uint64_t i);
uint64_t bit_rev_twice(unsigned nbits, unsigned base_nbits, uint64_t i);
uint64_t bit_rev_twice_inv(unsigned nbits, unsigned base_nbits, uint64_t i);

// four arrays of n doubles: re hi, re lo, im hi, im lo; entry 0 stays 0.0
void init_negacyclic_twiddles(size_t n, double *re0, double *re1, double *im0, double *im1);

} // namespace cfft

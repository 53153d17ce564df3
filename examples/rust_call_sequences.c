/* rust_call_sequences.c -- replays, call for call, what every method of the Rust crate rust/concrete-fft-b200 does
 * through the C ABI (the image has no rustc, so this is how the crate's call sequences are exercised): plan creation with
 * the reference's panics as statuses, queries, host transforms, clone / drop, fwd_monomial, the serde standard-order
 * mapping including its invalid_length paths (too short AND too long sequences still scatter the first n elements,
 * src/unordered.rs:1019-1031), autotune + report, kernel_name, fft128 fwd / inv, as_raw() + the device:: entry points on
 * device memory, and the polynomial host entry.  Exit code 0 = every sequence behaved; 77 = no CUDA device.
 *
 *   gcc -std=c99 -Wall -Iinclude examples/rust_call_sequences.c -Lconcrete_fft_b200 -lcfft_b200 \
 *       -Wl,-rpath,$PWD/concrete_fft_b200 -L/usr/local/cuda/lib64 -lcudart -lm -o rust_call_sequences
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cfft_b200.h"

/* the three runtime calls the device:: section needs; declared here so that the file stays plain C99 without CUDA headers */
extern int cudaMalloc(void **p, size_t bytes);
extern int cudaFree(void *p);
extern int cudaMemcpy(void *dst, const void *src, size_t bytes, int kind); /* 1 = host to device, 2 = device to host */
extern int cudaDeviceSynchronize(void);

#define CHECK(cond)                                                                          \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            fprintf(stderr, "FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #cond, cfft_last_error()); \
            return 1;                                                                        \
        }                                                                                    \
    } while (0)

static double frand(unsigned *s)
{
    *s = *s * 1664525u + 1013904223u;
    return (double)(*s >> 8) / 16777216.0;
}

int main(void)
{
    cfft_plan *p = NULL, *q = NULL;
    unsigned seed = 1;

    /* ---- ordered::Plan::new / panics (src/ordered.rs:242-244) ---- */
    cfft_status st = cfft_ordered_plan_create(&p, 0, 64, CFFT_METHOD_USER, CFFT_DIF4, 0);
    if (st == CFFT_ECUDA) {
        fprintf(stderr, "no CUDA device: %s\n", cfft_last_error());
        return 77;
    }
    CHECK(st == CFFT_OK);
    CHECK(cfft_ordered_plan_create(&q, 0, 48, CFFT_METHOD_USER, CFFT_DIF4, 0) == CFFT_EINVAL);   /* not a power of two */
    CHECK(cfft_ordered_plan_create(&q, 0, 2048, CFFT_METHOD_USER, CFFT_DIF4, 0) == CFFT_EINVAL); /* > 2^10 */
    /* fft_size / algo / fft_scratch */
    int algo = -1;
    uint64_t base = 0, bytes = 0, align = 0;
    CHECK(cfft_plan_fft_size(p) == 64);
    CHECK(cfft_plan_algo(p, &algo, &base) == CFFT_OK && algo == CFFT_DIF4 && base == 64);
    CHECK(cfft_plan_scratch_req(p, &bytes, &align) == CFFT_OK && bytes == 64 * 16 && align == 128);
    /* fwd / inv on a host slice (batch = 1), Clone, Drop */
    double x[128], y[128];
    for (int i = 0; i < 128; i++) x[i] = y[i] = frand(&seed);
    CHECK(cfft_c64_fwd_host(p, y, 64, 1) == CFFT_OK);
    CHECK(cfft_plan_clone(p, &q) == CFFT_OK);
    CHECK(cfft_c64_inv_host(q, y, 64, 1) == CFFT_OK);
    for (int i = 0; i < 128; i++) CHECK(fabs(y[i] / 64.0 - x[i]) < 1e-13);
    CHECK(cfft_c64_fwd_host(p, y, 63, 1) == CFFT_ELENGTH); /* the length assert */
    cfft_plan_destroy(q);
    cfft_plan_destroy(p);

    /* ---- unordered::Plan ---- */
    const uint64_t n = 1024;
    CHECK(cfft_unordered_plan_create(&p, 0, n, CFFT_METHOD_USER, CFFT_DIF16, 256) == CFFT_OK);
    CHECK(cfft_unordered_plan_create(&q, 0, n, CFFT_METHOD_USER, CFFT_DIF16, 16) == CFFT_EINVAL); /* base_n < 32 and != n */
    CHECK(cfft_plan_algo(p, &algo, &base) == CFFT_OK && algo == CFFT_DIF16 && base == 256);
    CHECK(cfft_plan_scratch_req(p, &bytes, &align) == CFFT_OK && bytes == 256 * 16);
    CHECK(strlen(cfft_plan_kernel_name(p)) > 0);
    /* Method::Measure: a second plan, autotuned on the device, same order as UserProvided{Dif16, 256} */
    CHECK(cfft_unordered_plan_create(&q, 0, n, CFFT_METHOD_MEASURE, 0, 0) == CFFT_OK);
    CHECK(cfft_plan_algo(q, &algo, &base) == CFFT_OK && algo == CFFT_DIF16 && base == 256);
    char report[4096];
    CHECK(cfft_plan_autotune(q, 0) == CFFT_OK);
    CHECK(cfft_plan_tuning_report(q, report, sizeof report) > 0 && strstr(report, "selected:") != NULL);
    cfft_plan_destroy(q);

    double *buf = malloc(sizeof(double) * 2 * n), *std_order = malloc(sizeof(double) * 2 * n), *back = malloc(sizeof(double) * 2 * n);
    uint64_t *perm = malloc(sizeof(uint64_t) * n);
    /* fwd_monomial(degree, buf) == fwd(X^degree) */
    CHECK(cfft_unordered_fwd_monomial_host(p, 5, buf, n) == CFFT_OK);
    memset(back, 0, sizeof(double) * 2 * n);
    back[2 * 5] = 1.0;
    CHECK(cfft_c64_fwd_host(p, back, n, 1) == CFFT_OK);
    for (uint64_t i = 0; i < 2 * n; i++) CHECK(fabs(buf[i] - back[i]) < 1e-12);
    CHECK(cfft_unordered_fwd_monomial_host(p, n, buf, n) == CFFT_EINVAL); /* degree < n */

    /* serialize_fourier_buffer: gather into standard order; deserialize: scatter element i to perm[i] */
    for (uint64_t i = 0; i < 2 * n; i++) buf[i] = frand(&seed);
    CHECK(cfft_unordered_permutation(p, perm) == CFFT_OK);
    CHECK(cfft_unordered_to_standard_host(p, buf, std_order) == CFFT_OK);
    for (uint64_t i = 0; i < n; i++) CHECK(std_order[2 * i] == buf[2 * perm[i]] && std_order[2 * i + 1] == buf[2 * perm[i] + 1]);
    memset(back, 0, sizeof(double) * 2 * n);
    CHECK(cfft_unordered_from_standard_host(p, std_order, n, back) == CFFT_OK);
    CHECK(memcmp(back, buf, sizeof(double) * 2 * n) == 0);
    /* invalid_length, too short: the n - 3 elements that arrived are in place, the status says so */
    memset(back, 0, sizeof(double) * 2 * n);
    CHECK(cfft_unordered_from_standard_host(p, std_order, n - 3, back) == CFFT_ELENGTH);
    for (uint64_t i = 0; i < n - 3; i++) CHECK(back[2 * perm[i]] == std_order[2 * i]);
    for (uint64_t i = n - 3; i < n; i++) CHECK(back[2 * perm[i]] == 0.0);
    /* invalid_length, too long: the first n elements are in place (the Rust visitor stops storing at n and keeps counting) */
    memset(back, 0, sizeof(double) * 2 * n);
    CHECK(cfft_unordered_from_standard_host(p, std_order, n + 2, back) == CFFT_ELENGTH);
    CHECK(memcmp(back, buf, sizeof(double) * 2 * n) == 0);

    /* ---- as_raw() + device::c64_fwd / c64_mul_assign / c64_inv / c64_fwd_mul_inv on device memory ---- */
    void *d_a = NULL, *d_b = NULL, *d_o = NULL;
    CHECK(cudaMalloc(&d_a, 16 * n) == 0 && cudaMalloc(&d_b, 16 * n) == 0 && cudaMalloc(&d_o, 16 * n) == 0);
    double *a = malloc(16 * n), *b = malloc(16 * n), *o1 = malloc(16 * n), *o2 = malloc(16 * n);
    for (uint64_t i = 0; i < 2 * n; i++) { a[i] = frand(&seed) - 0.5; b[i] = frand(&seed) - 0.5; }
    CHECK(cudaMemcpy(d_a, a, 16 * n, 1) == 0 && cudaMemcpy(d_b, b, 16 * n, 1) == 0);
    CHECK(cfft_c64_fwd(p, d_b, 1, NULL) == CFFT_OK);                               /* b to the Fourier domain */
    CHECK(cfft_c64_fwd_mul_inv(p, d_a, 1, d_b, 0, d_o, 1, NULL) == CFFT_OK);        /* fused: inv(fwd(a) * b) */
    CHECK(cudaMemcpy(o1, d_o, 16 * n, 2) == 0);
    CHECK(cfft_c64_fwd(p, d_a, 1, NULL) == CFFT_OK);                               /* the same as three calls */
    CHECK(cfft_c64_mul_assign(0, d_a, d_b, n, NULL) == CFFT_OK);
    CHECK(cfft_c64_inv(p, d_a, 1, NULL) == CFFT_OK);
    CHECK(cudaMemcpy(o2, d_a, 16 * n, 2) == 0);
    CHECK(memcmp(o1, o2, 16 * n) == 0);                                             /* bit-identical */
    CHECK(cfft_c64_fwd(p, (char *)d_a + 8, 1, NULL) == CFFT_EINVAL);                /* 16-byte alignment of device buffers */

    /* ---- device::c64_fwd_strided / c64_inv_strided: rows inside larger records; multi_gpu::Replicas ---- */
    {
        void *d_rec = NULL;
        double *rec = malloc(16 * 3 * 2 * n), *got = malloc(16 * 3 * 2 * n), *row = malloc(16 * n);
        for (uint64_t i = 0; i < 2 * 3 * 2 * n; i++) rec[i] = frand(&seed);
        CHECK(cudaMalloc(&d_rec, 16 * 3 * 2 * n) == 0 && cudaMemcpy(d_rec, rec, 16 * 3 * 2 * n, 1) == 0);
        /* three records of two polynomials each: transform the SECOND polynomial of every record where it is */
        CHECK(cfft_c64_fwd_strided(p, (char *)d_rec + 16 * n, 2 * n, 3, NULL) == CFFT_OK);
        CHECK(cudaMemcpy(got, d_rec, 16 * 3 * 2 * n, 2) == 0);
        for (int r = 0; r < 3; r++) {
            memcpy(row, rec + (size_t)(2 * r + 1) * 2 * n, 16 * n);
            CHECK(cfft_c64_fwd_host(p, row, n, 1) == CFFT_OK);
            CHECK(memcmp(row, got + (size_t)(2 * r + 1) * 2 * n, 16 * n) == 0);                         /* same bits as the packed call */
            CHECK(memcmp(rec + (size_t)(2 * r) * 2 * n, got + (size_t)(2 * r) * 2 * n, 16 * n) == 0);   /* neighbours untouched */
        }
        CHECK(cfft_c64_inv_strided(p, (char *)d_rec + 16 * n, 2 * n, 3, NULL) == CFFT_OK);
        CHECK(cfft_c64_fwd_strided(p, d_rec, n / 2, 3, NULL) == CFFT_EINVAL);                           /* rows would overlap */
        /* Replicas::new(plan.as_raw(), &[0, 0]) and Replicas::c64(2, buf): one host call over two replicas */
        cfft_plan *rep[2] = {NULL, NULL};
        const cfft_plan *crep[2];
        CHECK(cfft_plan_clone_to_device(p, 0, &rep[0]) == CFFT_OK && cfft_plan_clone_to_device(p, 0, &rep[1]) == CFFT_OK);
        CHECK(cfft_plan_clone_to_device(p, 9999, &q) == CFFT_EINVAL);
        crep[0] = rep[0];
        crep[1] = rep[1];
        memcpy(got, rec, 16 * 3 * 2 * n);
        CHECK(cfft_c64_host_multi(crep, 2, 2, got, 6 * n, 6) == CFFT_OK);                               /* fwd then inv of six rows */
        for (uint64_t i = 0; i < 2 * 6 * n; i++) CHECK(fabs(got[i] / (double)n - rec[i]) < 1e-11);
        CHECK(cfft_c64_host_multi(crep, 2, 0, got, 6 * n - 1, 6) == CFFT_ELENGTH);
        CHECK(cfft_c64_host_multi(crep, 0, 0, got, 6 * n, 6) == CFFT_EINVAL);
        cfft_plan_destroy(rep[0]);
        cfft_plan_destroy(rep[1]);
        cudaFree(d_rec);
        free(rec); free(got); free(row);
    }

    /* ---- the polynomial host entry: integers in host memory, operand resident on the device ---- */
    int64_t *pa = malloc(sizeof(int64_t) * 2 * n), *pb = malloc(sizeof(int64_t) * 2 * n), *pc = malloc(sizeof(int64_t) * 2 * n);
    void *d_pb = NULL;
    for (uint64_t i = 0; i < 2 * n; i++) { pa[i] = (int64_t)(frand(&seed) * 2000.0) - 1000; pb[i] = 0; }
    pb[1] = 1; /* b = X: the negacyclic product a * X rotates the coefficients by one and negates the wrapped one */
    CHECK(cudaMalloc(&d_pb, 16 * n) == 0 && cudaMemcpy(d_pb, pb, 16 * n, 1) == 0);
    CHECK(cfft_c64_poly_fwd(p, (const int64_t *)d_pb, d_b, 1, CFFT_POLY_INTEGER, NULL) == CFFT_OK);
    CHECK(cudaDeviceSynchronize() == 0);
    CHECK(cfft_c64_poly_mul_host(p, pa, 1, d_b, 0, pc, 1, CFFT_POLY_INTEGER) == CFFT_OK);
    CHECK(pc[0] == -pa[2 * n - 1]);
    for (uint64_t i = 1; i < 2 * n; i++) CHECK(pc[i] == pa[i - 1]);
    CHECK(cfft_c64_poly_mul_host(p, pa, 0, d_b, 0, pc, 1, CFFT_POLY_INTEGER) == CFFT_EINVAL); /* k_terms >= 1 */
    cfft_plan_destroy(p);

    /* ---- fft128::Plan ---- */
    CHECK(cfft_f128_plan_create(&p, 0, 16) == CFFT_EINVAL); /* n >= 32 */
    CHECK(cfft_f128_plan_create(&p, 0, 64) == CFFT_OK);
    double pl[4][64], orig[4][64];
    for (int k = 0; k < 4; k++)
        for (int i = 0; i < 64; i++) orig[k][i] = pl[k][i] = (k & 1) ? 0.0 : frand(&seed);
    CHECK(cfft_f128_fwd_host(p, pl[0], pl[1], pl[2], pl[3], 64, 1) == CFFT_OK);
    CHECK(cfft_f128_inv_host(p, pl[0], pl[1], pl[2], pl[3], 64, 1) == CFFT_OK);
    for (int i = 0; i < 64; i++) CHECK(fabs((pl[0][i] + pl[1][i]) / 64.0 - orig[0][i]) < 1e-28 && fabs((pl[2][i] + pl[3][i]) / 64.0 - orig[2][i]) < 1e-28);
    CHECK(cfft_f128_fwd_host(p, pl[0], pl[1], pl[2], pl[3], 63, 1) == CFFT_ELENGTH);
    CHECK(cfft_plan_algo(p, &algo, &base) == CFFT_EINVAL); /* no algo on an fft128 plan */
    cfft_plan_destroy(p);

    cudaFree(d_a); cudaFree(d_b); cudaFree(d_o); cudaFree(d_pb);
    free(buf); free(std_order); free(back); free(perm); free(a); free(b); free(o1); free(o2); free(pa); free(pb); free(pc);
    printf("rust_call_sequences: all sequences ok\n");
    return 0;
}

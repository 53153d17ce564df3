/* c_api_demo.c -- the C ABI used directly from compiled code (what the Rust crate's FFI calls do).
 *
 *   gcc -std=c99 -Wall -Iinclude examples/c_api_demo.c -Lconcrete_fft_b200 -lcfft_b200 \
 *       -Wl,-rpath,$PWD/concrete_fft_b200 -lm -o c_api_demo && ./c_api_demo
 *
 * unordered::Plan::new(2048, UserProvided{Dif16, 256}); fwd; inv on host memory; checks the round trip
 * and prints the first Fourier coefficient (position 0 holds X_0 = sum of the inputs in every plan). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cfft_b200.h"

int main(void)
{
    const uint64_t n = 2048, batch = 4;
    cfft_plan *plan = NULL;
    cfft_status st = cfft_unordered_plan_create(&plan, 0, n, CFFT_METHOD_USER, CFFT_DIF16, 256);
    if (st != CFFT_OK) {
        fprintf(stderr, "plan creation failed: %s (%s)\n", cfft_status_string(st), cfft_last_error());
        return st == CFFT_ECUDA ? 77 : 1; /* 77: no CUDA device -- there is no CPU fallback */
    }
    double *buf = malloc(sizeof(double) * 2 * n * batch), sum_re = 0.0;
    for (uint64_t i = 0; i < n * batch; i++) {
        buf[2 * i] = (double)(i % 17) / 17.0;
        buf[2 * i + 1] = (double)(i % 5) / 5.0;
        if (i < n) sum_re += buf[2 * i];
    }
    st = cfft_c64_fwd_host(plan, buf, n * batch, batch);
    if (st != CFFT_OK) { fprintf(stderr, "fwd: %s\n", cfft_last_error()); return 1; }
    printf("%s, kernel %s: X_0.re = %.17g (sum of inputs %.17g)\n", cfft_version(), cfft_plan_kernel_name(plan), buf[0], sum_re);
    st = cfft_c64_inv_host(plan, buf, n * batch, batch);
    if (st != CFFT_OK) { fprintf(stderr, "inv: %s\n", cfft_last_error()); return 1; }
    double err = 0.0;
    for (uint64_t i = 0; i < n * batch; i++) {
        err = fmax(err, fabs(buf[2 * i] / (double)n - (double)(i % 17) / 17.0));
        err = fmax(err, fabs(buf[2 * i + 1] / (double)n - (double)(i % 5) / 5.0));
    }
    printf("round trip max error %.3g\n", err);
    /* a wrong length is the assert_eq! of src/unordered.rs:827 */
    if (cfft_c64_fwd_host(plan, buf, n * batch - 1, batch) != CFFT_ELENGTH) return 1;
    cfft_plan_destroy(plan);
    free(buf);
    return err < 1e-12 ? 0 : 1;
}

/* parallel.c -- pthread row splitter for the oracle's batch drivers (test infrastructure). */
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>

struct span { void (*fn)(void *, size_t, size_t); void *ctx; size_t lo, hi; };

static void *span_main(void *arg)
{
    struct span *s = arg;
    s->fn(s->ctx, s->lo, s->hi);
    return NULL;
}

void orc_parallel_rows(int threads, size_t total, void (*fn)(void *ctx, size_t lo, size_t hi), void *ctx)
{
    if (threads <= 1 || total < 2) {
        fn(ctx, 0, total);
        return;
    }
    if ((size_t)threads > total)
        threads = (int)total;
    pthread_t *tid = malloc((size_t)threads * sizeof *tid);
    struct span *sp = malloc((size_t)threads * sizeof *sp);
    for (int t = 0; t < threads; t++) {
        sp[t].fn = fn;
        sp[t].ctx = ctx;
        sp[t].lo = total * (size_t)t / (size_t)threads;
        sp[t].hi = total * (size_t)(t + 1) / (size_t)threads;
        pthread_create(&tid[t], NULL, span_main, &sp[t]);
    }
    for (int t = 0; t < threads; t++)
        pthread_join(tid[t], NULL);
    free(tid);
    free(sp);
}

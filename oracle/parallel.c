/* parallel.c -- row splitter for the oracle's batch drivers (test infrastructure).
 *
 * A small persistent pthread pool (workers park on a condition variable between calls), so that
 * the timed CPU baseline does not pay thread creation per transform pass -- the way a Rust caller
 * would run the reference over a batch with a rayon-style pool.  One job at a time (calls are
 * serialised by a mutex); the calling thread works on span 0 itself. */
#include "oracle.h"

#include <pthread.h>
#include <stdlib.h>

#define ORC_MAX_THREADS 256

struct job {
    void (*fn)(void *, size_t, size_t);
    void *ctx;
    size_t total;
    int parts;
};

static pthread_mutex_t g_call = PTHREAD_MUTEX_INITIALIZER; /* one job at a time */
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_go = PTHREAD_COND_INITIALIZER, g_done = PTHREAD_COND_INITIALIZER;
static pthread_t g_tid[ORC_MAX_THREADS];
static int g_workers = 0;        /* threads created so far (worker w serves span w + 1) */
static unsigned long g_epoch = 0; /* bumped per job */
static int g_pending = 0;        /* workers still running the current job */
static struct job g_job;

static void run_span(const struct job *j, int part)
{
    const size_t lo = j->total * (size_t)part / (size_t)j->parts;
    const size_t hi = j->total * (size_t)(part + 1) / (size_t)j->parts;
    if (hi > lo)
        j->fn(j->ctx, lo, hi);
}

static void *worker_main(void *arg)
{
    const int part = (int)(size_t)arg; /* span index served by this worker */
    unsigned long seen = 0;
    pthread_mutex_lock(&g_mu);
    for (;;) {
        while (g_epoch == seen)
            pthread_cond_wait(&g_go, &g_mu);
        seen = g_epoch;
        if (part < g_job.parts) {
            const struct job j = g_job;
            pthread_mutex_unlock(&g_mu);
            run_span(&j, part);
            pthread_mutex_lock(&g_mu);
            if (--g_pending == 0)
                pthread_cond_signal(&g_done);
        }
    }
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
    if (threads > ORC_MAX_THREADS)
        threads = ORC_MAX_THREADS;

    pthread_mutex_lock(&g_call);
    pthread_mutex_lock(&g_mu);
    while (g_workers < threads - 1) { /* grow the pool on demand; new workers start with seen = 0 < epoch */
        if (pthread_create(&g_tid[g_workers], NULL, worker_main, (void *)(size_t)(g_workers + 1)) != 0)
            break;
        pthread_detach(g_tid[g_workers]);
        g_workers++;
    }
    int parts = g_workers + 1 < threads ? g_workers + 1 : threads;
    g_job.fn = fn;
    g_job.ctx = ctx;
    g_job.total = total;
    g_job.parts = parts;
    g_pending = parts - 1;
    g_epoch++;
    pthread_cond_broadcast(&g_go);
    pthread_mutex_unlock(&g_mu);

    run_span(&g_job, 0);

    pthread_mutex_lock(&g_mu);
    while (g_pending > 0)
        pthread_cond_wait(&g_done, &g_mu);
    pthread_mutex_unlock(&g_mu);
    pthread_mutex_unlock(&g_call);
}

/*
 * bvg_oracle_mt.c -- multi-threaded driver over the oracle's sequential scan.  TEST/BENCH
 * INFRASTRUCTURE ONLY (see bvg_oracle.h).
 *
 * Splits [from, to) into `threads` node ranges of ceil(len/threads) nodes exactly as
 * ImmutableGraph.splitNodeIterators does (reference src/it/unimi/dsi/webgraph/ImmutableGraph.java:
 * 379-409): each range is an independent nodeIterator(from_i) whose window is seeded by random
 * access (BVGraph.java:1173-1183).  Used as the all-cores CPU baseline.
 */
#include "bvg_oracle.h"
#include <pthread.h>
#include <stdlib.h>

typedef struct { const orc_graph* g; int32_t from, to; int64_t arcs; uint64_t cs; int rc; } job_t;

static void* run(void* p) {
    job_t* j = (job_t*)p;
    j->rc = orc_scan_range(j->g, j->from, j->to, &j->arcs, &j->cs);
    return NULL;
}

int orc_scan_range_mt(const orc_graph* g, int32_t from, int32_t to, int threads,
                      int64_t* arcs, uint64_t* checksum) {
    if (threads < 1) threads = 1;
    if (from < 0 || to > g->n || to < from) return BVGO_EINVAL;
    const int64_t len = (int64_t)to - from;
    const int64_t step = (len + threads - 1) / threads;
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    int used = 0;
    for (int i = 0; i < threads; i++) {
        int64_t a = from + i * step, b = a + step;
        if (a >= to) break;
        if (b > to) b = to;
        jobs[i].g = g; jobs[i].from = (int32_t)a; jobs[i].to = (int32_t)b;
        pthread_create(&th[i], NULL, run, &jobs[i]);
        used++;
    }
    int rc = BVGO_OK;
    int64_t tot = 0;
    uint64_t cs = 0;
    for (int i = 0; i < used; i++) {
        pthread_join(th[i], NULL);
        if (jobs[i].rc < 0 && rc == BVGO_OK) rc = jobs[i].rc;
        tot += jobs[i].arcs;
        cs ^= jobs[i].cs;
    }
    free(jobs); free(th);
    if (rc == BVGO_OK) { *arcs = tot; *checksum = cs; }
    return rc;
}

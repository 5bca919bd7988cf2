/*
 * katgpu_double.c — a TEST DOUBLE of libkatgpu for machines without a GPU.
 *
 * TEST INFRASTRUCTURE ONLY (tests/test_host_frontend.py).  It exports the part of include/katgpu.h the
 * C++ front-end (kat_b200/bin/kat) calls and answers every call with the CPU oracle
 * (oracle/libkatoracle.so), so that the front-end's HOST logic — option handling, input groups,
 * the FASTA/FASTQ readers, LOAD mode, the text and jellyfish-hash writers, sect/cold/filter
 * arithmetic — is checked against the reference-written golden files in the `-m "not gpu"` suite.
 * It is never built into, shipped with or found by the product: kat_b200/bin/kat links and loads the real
 * libkatgpu.so (RUNPATH $ORIGIN/../lib); the tests interpose THIS library's symbols with LD_PRELOAD for the
 * processes they start.  What the GPU computes is checked by the `-m gpu` tests, which run the same
 * commands against the real library.
 *
 * Several devices (KATGPU_DOUBLE_DEVICES=N, default 1): `kat --gpus N` then runs its N host threads against N contexts of
 * the double, joined into one job by katgpu_dist_init.  A job's calls are collective, like the real library's: the i-th
 * table begun on every rank is ONE shared oracle table, every rank's feed adds its share to it, katgpu_count_end waits for
 * all ranks, walks and lookups give every rank the job's (all-reduced) answer, and exports hand each rank its own part
 * of the k-mers.  That checks the front-end's sharding of a base stream over the GPUs (kat_front.hpp::feed_all: start
 * positions split N ways, k-1 bases of overlap), per-GPU table sizing and the merging of per-rank parts.
 *
 * Built by tests/hostcheck/Makefile into tests/hostcheck/_build/libkatgpu_double.so (git-ignored, gpurun-ignored).
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "katgpu.h"

/* oracle/kat_oracle.c */
typedef struct oracle_table oracle_table;
oracle_table *oracle_table_new(int k, int canonical);
void oracle_table_free(oracle_table *t);
void oracle_table_add(oracle_table *t, const unsigned char *bases, size_t n);
void oracle_table_add_pairs(oracle_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, size_t n);
void oracle_table_finish(oracle_table *t);
uint64_t oracle_table_distinct(const oracle_table *t);
uint64_t oracle_table_total(const oracle_table *t);
void oracle_table_export_wide(const oracle_table *t, uint64_t *keys_hi, uint64_t *keys_lo, uint64_t *counts);
uint64_t oracle_lookup_wide(const oracle_table *t, uint64_t key_hi, uint64_t key_lo, int canonicalise);
void oracle_hist(const oracle_table *t, uint64_t base, uint64_t ceil_, uint64_t inc, uint64_t nb, uint64_t *bins);
void oracle_gcp(const oracle_table *t, double cvg_scale, uint32_t cvg_bins, uint64_t *mx);
void oracle_comp(const oracle_table *t1, const oracle_table *t2, const oracle_table *t3, double d1_scale, double d2_scale,
                 uint32_t d1_bins, uint32_t d2_bins, uint64_t *main_mx, uint64_t *ends_mx, uint64_t *middle_mx, uint64_t *mixed_mx,
                 uint64_t *counters, uint64_t *spectra);

/* one oracle table, shared by the ranks of a job (refs = handles alive) */
typedef struct {
    oracle_table *o;
    int ended, refs;
} shared_table;
#define MAX_JOB_TABLES 64
typedef struct {
    int n_ranks;
    pthread_mutex_t mu;
    pthread_barrier_t bar;
    shared_table *tables[MAX_JOB_TABLES]; /* the i-th table begun on each rank */
} job_t;
static job_t g_job; /* the front-end runs one job per process */
static pthread_mutex_t g_job_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_job_ranks_joined = 0;

struct katgpu_ctx {
    job_t *job; /* NULL: a single device */
    int rank, n_begun;
};
struct katgpu_table {
    shared_table *sh;
    oracle_table *o; /* = sh->o */
    katgpu_ctx *ctx;
    int k, canonical, allow_grow;
    uint64_t slots;
};

static __thread char g_err[256];
static katgpu_status fail(katgpu_status st, const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return st;
}
const char *katgpu_last_error(void) { return g_err; }
int katgpu_version(void) { return KATGPU_VERSION; }
int katgpu_device_count(void) {
    const char *e = getenv("KATGPU_DOUBLE_DEVICES");
    const int n = e ? atoi(e) : 1;
    return n < 1 ? 1 : n > 8 ? 8 : n;
}
katgpu_status katgpu_init(int device, katgpu_ctx **out) {
    if (device < 0 || device >= katgpu_device_count()) return fail(KATGPU_ERR_NO_DEVICE, "no such device");
    *out = (katgpu_ctx *)calloc(1, sizeof(katgpu_ctx));
    return KATGPU_OK;
}
void katgpu_shutdown(katgpu_ctx *ctx) { free(ctx); }
katgpu_status katgpu_dist_unique_id(void *id128) { memset(id128, 0, 128); return KATGPU_OK; }
katgpu_status katgpu_dist_init(katgpu_ctx *ctx, int rank, int n_ranks, const void *id128) {
    (void)id128;
    if (n_ranks < 2 || n_ranks > katgpu_device_count() || rank < 0 || rank >= n_ranks) return fail(KATGPU_ERR_INVALID, "the test double has fewer devices");
    pthread_mutex_lock(&g_job_mu);
    if (g_job_ranks_joined++ == 0) {
        memset(&g_job, 0, sizeof g_job);
        g_job.n_ranks = n_ranks;
        pthread_mutex_init(&g_job.mu, NULL);
        pthread_barrier_init(&g_job.bar, NULL, (unsigned)n_ranks);
    }
    pthread_mutex_unlock(&g_job_mu);
    ctx->job = &g_job;
    ctx->rank = rank;
    return KATGPU_OK;
}

static void lock_job(katgpu_table *t) { if (t->ctx->job) pthread_mutex_lock(&t->ctx->job->mu); }
static void unlock_job(katgpu_table *t) { if (t->ctx->job) pthread_mutex_unlock(&t->ctx->job->mu); }

katgpu_status katgpu_count_begin(katgpu_ctx *ctx, int k, int canonical, uint64_t hash_size, int allow_grow, katgpu_table **out) {
    if (k < 1 || k > 64) return fail(KATGPU_ERR_INVALID, "k out of 1..64");
    if (k > 32 && ctx->job) return fail(KATGPU_ERR_INVALID, "k > 32 is counted on one GPU (not on a multi-GPU context)"); /* as the real library */
    katgpu_table *t = (katgpu_table *)calloc(1, sizeof *t);
    if (ctx->job) { /* collective: the i-th begin of every rank names the same table */
        job_t *j = ctx->job;
        const int i = ctx->n_begun++;
        if (i >= MAX_JOB_TABLES) return fail(KATGPU_ERR_INVALID, "too many tables in one job of the test double");
        pthread_mutex_lock(&j->mu);
        if (!j->tables[i]) {
            j->tables[i] = (shared_table *)calloc(1, sizeof(shared_table));
            j->tables[i]->o = oracle_table_new(k, canonical);
        }
        t->sh = j->tables[i];
        t->sh->refs++;
        pthread_mutex_unlock(&j->mu);
    } else {
        t->sh = (shared_table *)calloc(1, sizeof(shared_table));
        t->sh->o = oracle_table_new(k, canonical);
        t->sh->refs = 1;
    }
    t->o = t->sh->o;
    t->ctx = ctx;
    t->k = k, t->canonical = canonical, t->allow_grow = allow_grow;
    t->slots = 1024;
    while (t->slots < hash_size) t->slots <<= 1;
    *out = t;
    return KATGPU_OK;
}
katgpu_status katgpu_count_feed(katgpu_table *t, const char *bases, size_t n) {
    if (t->sh->ended) return fail(KATGPU_ERR_STATE, "feed after end");
    lock_job(t);
    oracle_table_add(t->o, (const unsigned char *)bases, n); /* every rank's share is a chunk of its own: no k-mer spans two */
    unlock_job(t);
    return KATGPU_OK;
}
/* Raw FASTQ records.  By default the double calls every block "not in the regular layout", so the front-end's host route runs.
 * With KATGPU_DOUBLE_FASTQ=1 it takes blocks the way the real library does (katgpu_fastq.cu, parse_kernels.cu: every line
 * '\n'-terminated, four lines per record, '@' header, sequence longer than the trim that neither starts with '@' / '+' nor
 * holds a '+', a '+' line, quality of the sequence's length) and counts the trimmed sequences, one chunk per record —
 * all or nothing per call — so that the front-end's raw route (blocks of whole records, per-GPU shares, starting an input
 * over on the host route when a later block is refused) is checked on the CPU as well. */
static const char *line_end(const char *p, const char *e) { return p < e ? (const char *)memchr(p, '\n', (size_t)(e - p)) : NULL; }
katgpu_status katgpu_count_feed_fastq(katgpu_table *t, const char *bytes, size_t n, unsigned trim5p, int *regular) {
    *regular = 0;
    if (!getenv("KATGPU_DOUBLE_FASTQ")) return KATGPU_OK;
    if (n == 0) {
        *regular = 1;
        return KATGPU_OK;
    }
    if (bytes[n - 1] != '\n') return KATGPU_OK;
    for (int emit = 0; emit < 2; emit++) { /* validate everything, then count */
        const char *p = bytes, *e = bytes + n;
        while (p < e) {
            const char *h = line_end(p, e), *s = h ? line_end(h + 1, e) : NULL, *pl = s ? line_end(s + 1, e) : NULL, *q = pl ? line_end(pl + 1, e) : NULL;
            if (!q) return KATGPU_OK;
            const char *seq = h + 1;
            const size_t slen = (size_t)(s - seq), qlen = (size_t)(q - (pl + 1));
            if (*p != '@' || s[1] != '+' || slen <= trim5p || slen != qlen || *seq == '@' || *seq == '+' || memchr(seq, '+', slen)) return KATGPU_OK;
            if (emit) {
                lock_job(t);
                oracle_table_add(t->o, (const unsigned char *)seq + trim5p, slen - trim5p);
                unlock_job(t);
            }
            p = q + 1;
        }
    }
    *regular = 1;
    return KATGPU_OK;
}
katgpu_status katgpu_count_reserve(katgpu_table *t, uint64_t expected_kmers) { (void)t; (void)expected_kmers; return KATGPU_OK; }
katgpu_status katgpu_count_add_pairs(katgpu_table *t, const uint64_t *keys, const uint64_t *counts, size_t n) {
    if (t->k > 32) return fail(KATGPU_ERR_INVALID, "k > 32: use katgpu_count_add_pairs_wide");
    lock_job(t);
    oracle_table_add_pairs(t->o, NULL, keys, counts, n);
    unlock_job(t);
    return KATGPU_OK;
}
katgpu_status katgpu_count_add_pairs_wide(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, const uint64_t *counts, size_t n) {
    lock_job(t);
    oracle_table_add_pairs(t->o, keys_hi, keys_lo, counts, n);
    unlock_job(t);
    return KATGPU_OK;
}
katgpu_status katgpu_count_end(katgpu_table *t, uint64_t *n_distinct, uint64_t *n_total) {
    if (t->ctx->job) pthread_barrier_wait(&t->ctx->job->bar); /* collective: every rank has fed its share */
    lock_job(t);
    if (!t->sh->ended) oracle_table_finish(t->o);
    t->sh->ended = 1;
    unlock_job(t);
    /* a job shares the k-mers out: each rank's table holds about 1/N of the distinct ones */
    const uint64_t mine = oracle_table_distinct(t->o) / (uint64_t)(t->ctx->job ? t->ctx->job->n_ranks : 1);
    if (!t->allow_grow && mine > t->slots) return fail(KATGPU_ERR_HASH_FULL, "Hash full");
    if (n_distinct) *n_distinct = oracle_table_distinct(t->o);
    if (n_total) *n_total = oracle_table_total(t->o);
    return KATGPU_OK;
}
void katgpu_table_free(katgpu_table *t) {
    if (!t) return;
    lock_job(t);
    const int left = --t->sh->refs;
    unlock_job(t);
    if (left == 0) { /* (a job's slot keeps the pointer: tables are not begun again after a free within one tool run) */
        oracle_table_free(t->sh->o);
        t->sh->o = NULL;
    }
    free(t);
}

katgpu_status katgpu_hist(katgpu_table *t, uint64_t base, uint64_t ceil, uint64_t inc, uint64_t nb_buckets, uint64_t *bins_out) {
    oracle_hist(t->o, base, ceil, inc, nb_buckets, bins_out);
    return KATGPU_OK;
}
katgpu_status katgpu_gcp(katgpu_table *t, double cvg_scale, uint32_t cvg_bins, uint64_t *mx_out) {
    oracle_gcp(t->o, cvg_scale, cvg_bins, mx_out);
    return KATGPU_OK;
}
katgpu_status katgpu_comp(katgpu_table *t1, katgpu_table *t2, katgpu_table *t3, double d1_scale, double d2_scale, uint32_t d1_bins,
                          uint32_t d2_bins, uint64_t *main_mx, uint64_t *ends_mx, uint64_t *middle_mx, uint64_t *mixed_mx,
                          katgpu_comp_counters *counters, uint64_t *spectra) {
    oracle_comp(t1->o, t2->o, t3 ? t3->o : NULL, d1_scale, d2_scale, d1_bins, d2_bins, main_mx, ends_mx, middle_mx, mixed_mx,
                (uint64_t *)counters, spectra);
    return KATGPU_OK;
}
katgpu_status katgpu_lookup_wide(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, size_t n, int canonicalise,
                                 uint64_t *counts_out) {
    for (size_t i = 0; i < n; i++) counts_out[i] = oracle_lookup_wide(t->o, keys_hi ? keys_hi[i] : 0, keys_lo[i], canonicalise);
    return KATGPU_OK;
}
katgpu_status katgpu_lookup(katgpu_table *t, const uint64_t *keys, size_t n, int canonicalise, uint64_t *counts_out) {
    if (t->k > 32) return fail(KATGPU_ERR_INVALID, "k > 32: use katgpu_lookup_wide");
    return katgpu_lookup_wide(t, NULL, keys, n, canonicalise, counts_out);
}
katgpu_status katgpu_table_export_wide(katgpu_table *t, uint64_t *keys_hi_out, uint64_t *keys_lo_out, uint64_t *counts_out, size_t cap,
                                       uint64_t *n_out) {
    uint64_t n = oracle_table_distinct(t->o);
    if (t->ctx->job) return fail(KATGPU_ERR_INVALID, "the test double exports a job's k-mers through katgpu_table_export_filtered_wide only");
    if (n_out) *n_out = n;
    if (cap < n) return KATGPU_OK; /* the real library writes the first `cap`; the front-end always sizes for all */
    uint64_t *hi = keys_hi_out ? keys_hi_out : (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    oracle_table_export_wide(t->o, hi, keys_lo_out, counts_out);
    /* "unordered": hand the records back in a scrambled but deterministic order so that nothing depends on sortedness */
    for (uint64_t i = 0, j = n; j > 1; i++, j--) {
        const uint64_t r = (i * 0x9E3779B97F4A7C15ULL >> 17) % j, a = j - 1;
        uint64_t x;
        x = hi[a], hi[a] = hi[r], hi[r] = x;
        x = keys_lo_out[a], keys_lo_out[a] = keys_lo_out[r], keys_lo_out[r] = x;
        x = counts_out[a], counts_out[a] = counts_out[r], counts_out[r] = x;
    }
    if (!keys_hi_out) free(hi);
    return KATGPU_OK;
}
katgpu_status katgpu_table_export(katgpu_table *t, uint64_t *keys_out, uint64_t *counts_out, size_t cap, uint64_t *n_out) {
    if (t->k > 32) return fail(KATGPU_ERR_INVALID, "k > 32: use katgpu_table_export_wide");
    return katgpu_table_export_wide(t, NULL, keys_out, counts_out, cap, n_out);
}

static unsigned gc_of(uint64_t hi, uint64_t lo) {
    return (unsigned)(__builtin_popcountll((lo ^ (lo >> 1)) & 0x5555555555555555ULL) + __builtin_popcountll((hi ^ (hi >> 1)) & 0x5555555555555555ULL));
}
katgpu_status katgpu_table_export_filtered_wide(katgpu_table *t, uint64_t low_count, uint64_t high_count, uint32_t low_gc, uint32_t high_gc,
                                                int inside, uint64_t *keys_hi_out, uint64_t *keys_lo_out, uint64_t *counts_out, size_t cap,
                                                uint64_t *n_out) {
    const uint64_t n = oracle_table_distinct(t->o);
    uint64_t *hi = (uint64_t *)malloc((n ? n : 1) * 8), *lo = (uint64_t *)malloc((n ? n : 1) * 8), *cn = (uint64_t *)malloc((n ? n : 1) * 8);
    uint64_t m = 0;
    oracle_table_export_wide(t->o, hi, lo, cn);
    const uint64_t ranks = t->ctx->job ? (uint64_t)t->ctx->job->n_ranks : 1, rank = t->ctx->job ? (uint64_t)t->ctx->rank : 0;
    for (uint64_t i = 0; i < n; i++) {
        if ((lo[i] * 0x9E3779B97F4A7C15ULL >> 40) % ranks != rank) continue; /* on a job every rank owns a share of the k-mers */
        const unsigned gc = gc_of(hi[i], lo[i]);
        const int in_bounds = low_gc <= gc && gc <= high_gc && low_count <= cn[i] && cn[i] <= high_count;
        if (in_bounds == (inside != 0)) {
            if (m < cap) {
                if (keys_hi_out) keys_hi_out[m] = hi[i];
                keys_lo_out[m] = lo[i], counts_out[m] = cn[i];
            }
            m++;
        }
    }
    free(hi); free(lo); free(cn);
    *n_out = m;
    return KATGPU_OK;
}
katgpu_status katgpu_table_export_filtered(katgpu_table *t, uint64_t low_count, uint64_t high_count, uint32_t low_gc, uint32_t high_gc,
                                           int inside, uint64_t *keys_out, uint64_t *counts_out, size_t cap, uint64_t *n_out) {
    if (t->k > 32) return fail(KATGPU_ERR_INVALID, "k > 32: use katgpu_table_export_filtered_wide");
    return katgpu_table_export_filtered_wide(t, low_count, high_count, low_gc, high_gc, inside, NULL, keys_out, counts_out, cap, n_out);
}

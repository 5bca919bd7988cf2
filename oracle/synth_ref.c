/*
 * synth_ref.c — benchmark tooling that lives beside the oracle (TEST / BENCH
 * INFRASTRUCTURE ONLY; the product never links or loads it):
 *
 *   synthref_reads()       the deterministic synthetic read generator of
 *                          SURVEY.md §8(d), restated independently of
 *                          libkatgpu (kat_b200/csrc/synth.cuh), so that
 *                          `bench.py --impl reference` can build its sample
 *                          without loading the product library.  tests check
 *                          it byte-equal to katgpu_synth_reads_host.
 *   synthref_instances()   number of k-mer instances a base stream holds
 *                          (windows of k consecutive ACGTacgt bytes) — the
 *                          invariant Σ(label·bins) of a `kat hist` output
 *                          must meet; used to reject runs of the reference in
 *                          which its unlocked `threadedData.push_back`
 *                          (src/histogram.cc:198) lost a thread's bins.
 *   synthref_write_fastq() base stream (read,'N',read,'N',…) -> 4-line FASTQ.
 *
 * Multi-threaded with pthreads: generating a 4 M-read sample takes < 1 s.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
static uint32_t genome_base(uint64_t seed_g, uint64_t pos) {
    const uint64_t w = splitmix64(seed_g * 0xD1342543DE82EF95ULL + (pos >> 5));
    return (uint32_t)(w >> (2 * (pos & 31))) & 3u;
}

typedef struct {
    uint64_t genome_len, seed_g, seed_r, first_read, lo, hi;
    uint32_t read_len, thr_sub, thr_n;
    char *out;
} job_t;

static void *synth_worker(void *arg) {
    const job_t *j = (const job_t *)arg;
    const uint32_t L = j->read_len;
    for (uint64_t r = j->lo; r < j->hi; r++) {
        const uint64_t idx = j->first_read + r;
        const uint64_t r0 = splitmix64(j->seed_r ^ splitmix64(idx));
        const uint64_t pos = (uint64_t)(((unsigned __int128)r0 * (j->genome_len - L + 1)) >> 64);
        const int rev = (int)(splitmix64(r0 ^ 0x5851F42D4C957F2DULL) & 1);
        char *o = j->out + r * ((uint64_t)L + 1);
        for (uint32_t b = 0; b < L; b++) {
            uint32_t c = rev ? 3u - genome_base(j->seed_g, pos + (L - 1 - b)) : genome_base(j->seed_g, pos + b);
            const uint64_t e = splitmix64(r0 + 0x632BE59BD9B4E019ULL * (uint64_t)(b + 1));
            if ((uint32_t)(e >> 40) < j->thr_n) {
                o[b] = 'N';
                continue;
            }
            if ((uint32_t)((e >> 16) & 0xFFFFFFu) < j->thr_sub) c = (c + 1u + (uint32_t)(e & 0xFFFFu) % 3u) & 3u;
            o[b] = "ACGT"[c];
        }
        o[L] = 'N';
    }
    return NULL;
}

/* out: n_reads * (read_len + 1) bytes.  Returns 0, or -1 on bad parameters. */
int synthref_reads(uint64_t genome_len, uint64_t seed_g, uint64_t seed_r, uint64_t first_read, uint64_t n_reads,
                   uint32_t read_len, double sub_rate, double n_rate, int threads, char *out) {
    if (!out || read_len == 0 || genome_len < read_len || sub_rate < 0 || sub_rate > 1 || n_rate < 0 || n_rate > 1) return -1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < threads; t++) {
        job_t *j = &jobs[t];
        j->genome_len = genome_len, j->seed_g = seed_g, j->seed_r = seed_r, j->first_read = first_read;
        j->lo = n_reads * (uint64_t)t / (uint64_t)threads;
        j->hi = n_reads * (uint64_t)(t + 1) / (uint64_t)threads;
        j->read_len = read_len;
        j->thr_sub = (uint32_t)(sub_rate * 16777216.0);
        j->thr_n = (uint32_t)(n_rate * 16777216.0);
        j->out = out;
        if (pthread_create(&th[t], NULL, synth_worker, j) != 0) {
            synth_worker(j);
            th[t] = 0;
        }
    }
    for (int t = 0; t < threads; t++)
        if (th[t]) pthread_join(th[t], NULL);
    return 0;
}

/* Windows of k consecutive bases (JFI/mer_dna.hpp:46-63 code table: ACGTacgt). */
uint64_t synthref_instances(const unsigned char *s, uint64_t n, int k) {
    uint64_t run = 0, total = 0;
    for (uint64_t i = 0; i < n; i++) {
        const unsigned char c = s[i] & 0xDF;
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
            if (++run >= (uint64_t)k) total++;
        } else
            run = 0;
    }
    return total;
}

/* stream = n_reads records of read_len bases + 1 separator.  Header "@r", quality 'I'. */
int synthref_write_fastq(const char *stream, uint64_t n_reads, uint32_t read_len, const char *path) {
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    const size_t rec = 3 + (size_t)read_len + 3 + (size_t)read_len + 1, batch = 4096;
    char *buf = (char *)malloc(rec * batch);
    if (!buf) {
        fclose(f);
        return -1;
    }
    int rc = 0;
    for (uint64_t r0 = 0; r0 < n_reads && rc == 0; r0 += batch) {
        const uint64_t nb = n_reads - r0 < batch ? n_reads - r0 : batch;
        for (uint64_t i = 0; i < nb; i++) {
            char *o = buf + i * rec;
            memcpy(o, "@r\n", 3);
            memcpy(o + 3, stream + (r0 + i) * ((uint64_t)read_len + 1), read_len);
            memcpy(o + 3 + read_len, "\n+\n", 3);
            memset(o + 6 + read_len, 'I', read_len);
            o[rec - 1] = '\n';
        }
        if (fwrite(buf, rec, nb, f) != nb) rc = -1;
    }
    free(buf);
    if (fclose(f) != 0) rc = -1;
    return rc;
}

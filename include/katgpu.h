/*
 * katgpu.h — C-ABI of libkatgpu, the Blackwell (sm_100a) k-mer spectra engine
 * that replaces the counting + histogram hot path behind `kat hist`,
 * `kat gcp` and `kat comp` (TGAC/KAT 2.4.2).
 *
 * The reference has no FFI for this path: the seam is the C++ class boundary
 * between libkat's InputHandler/JellyfishHelper (producer of the hash) and the
 * three tools' *Slice workers (consumers).  Each entry point below names the
 * reference interface it replaces (file:line relative to the KAT tree;
 * JFI = deps/jellyfish-2.2.0/include/jellyfish).  INTEGRATION.md shows the
 * binding a KAT maintainer would add.
 *
 * Conventions: plain C; opaque handles; every call returns a katgpu_status
 * (0 = ok, < 0 = error) and leaves a human-readable message for
 * katgpu_last_error() (thread-local) — no exception crosses the ABI.  The
 * caller owns all output buffers (row-major uint64_t, host memory unless the
 * name says _device).  The library owns device memory and streams.  One
 * in-flight call per context; one context per GPU per process.  There is NO
 * CPU fallback: without a CUDA device katgpu_init fails with
 * KATGPU_ERR_NO_DEVICE.
 */
#ifndef KATGPU_H
#define KATGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KATGPU_VERSION 100 /* 0.1.0 */

typedef int katgpu_status;
enum {
    KATGPU_OK = 0,
    KATGPU_ERR_INVALID = -1,   /* bad argument (k out of 1..64, null pointer, ...)      */
    KATGPU_ERR_CUDA = -2,      /* CUDA runtime error, text in katgpu_last_error()       */
    KATGPU_ERR_HASH_FULL = -3, /* table full and growth disabled: the reference's
                                  std::runtime_error("Hash full"), JFI/hash_counter.hpp:199 */
    KATGPU_ERR_NOMEM = -4,     /* device or host allocation failed                      */
    KATGPU_ERR_NO_DEVICE = -5, /* no CUDA device: the product has no CPU path           */
    KATGPU_ERR_STATE = -6,     /* call made in the wrong state (feed after end, ...)    */
    KATGPU_ERR_OVERFLOW = -7   /* internal side table exhausted (count-overflow keys)   */
};

typedef struct katgpu_ctx katgpu_ctx;     /* one GPU: device, streams, staging buffers */
typedef struct katgpu_table katgpu_table; /* one counted input: replaces the pair
                                             HashCounter + LargeHashArrayPtr held by
                                             InputHandler (lib/include/kat/input_handler.hpp:53-55) */

/* Last error text of the calling thread ("" if none). */
const char *katgpu_last_error(void);
int katgpu_version(void);
/* Number of visible CUDA devices (0 when there is no driver/GPU). */
int katgpu_device_count(void);

/* Context on CUDA device `device` (ordinal as CUDA sees it). */
katgpu_status katgpu_init(int device, katgpu_ctx **out);
void katgpu_shutdown(katgpu_ctx *ctx);

/* Pinned host memory for callers that want full PCIe rate on katgpu_count_feed. */
katgpu_status katgpu_host_alloc(size_t bytes, void **out);
void katgpu_host_free(void *p);

/* ---- counting: replaces InputHandler::count (lib/src/input_handler.cc:180-202),
 * i.e. `new HashCounter(hashSize, 2k, 7, threads)` + JellyfishHelper::countSeqFile /
 * countSlice (lib/src/jellyfish_helper.cc:219-246, 202-211) with its
 * mer_iterator (JFI/mer_iterator.hpp:59-89) and hash_counter::add
 * (JFI/hash_counter.hpp:98-130).
 *
 *   k              mer length, 1..64 (mer_dna::k, jellyfish_helper.cc:229).  k <= 32: one 64-bit word per
 *                  key; k = 33..64: two words ("wide" tables, single GPU; see the _wide entry points)
 *   canonical      count min(m, revcomp(m)) (InputHandler::canonical)
 *   hash_size      initial number of table slots, rounded up to a power of two
 *                  (InputHandler::hashSize, `-H`, default 1e8)
 *   allow_grow     0 = the reference's --disable_hash_grow: a full table is
 *                  KATGPU_ERR_HASH_FULL; 1 = double and rehash
 *                  (hash_counter::double_size, JFI/hash_counter.hpp:204-244)
 */
katgpu_status katgpu_count_begin(katgpu_ctx *ctx, int k, int canonical, uint64_t hash_size,
                                 int allow_grow, katgpu_table **out);

/* Feed `n` bytes of base stream — exactly what the reference parser hands to
 * mer_iterator (JFI/mer_overlap_sequence_parser.hpp:189-246): sequence
 * characters with one 'N' between records.  A/a C/c G/g T/t are bases; every
 * other byte breaks the k-mer run (JFI/mer_dna.hpp:46-63).  No k-mer spans two
 * calls (the reference never spans two files, :149).  `bases` is HOST memory;
 * the host->device copies are pipelined inside the call. */
katgpu_status katgpu_count_feed(katgpu_table *t, const char *bases, size_t n);
/* Same, for a stream already resident in device memory (16-byte aligned). */
katgpu_status katgpu_count_feed_device(katgpu_table *t, const void *bases_dev, size_t n);
/* Feed RAW FILE BYTES: `n` bytes (at most 1 GiB per call, HOST memory, pageable is fine) holding
 * WHOLE records of a 4-line FASTQ file.  The bytes are copied to the GPU and parsed THERE
 * (JFI/mer_overlap_sequence_parser.hpp:218-289 read_fastq / read_sequence / skip_quals for the
 * regular layout): newline index, per-record validation ('@' header, sequence longer than trim5p
 * that neither starts with '@'/'+' nor contains '+', '+' line, quality of the same length, every
 * line '\n'-terminated), then every non-sequence byte and the first trim5p bases of every read are
 * blanked in place and the result is counted like katgpu_count_feed_device.
 *   *regular = 1  the bytes were in the regular layout and have been counted;
 *   *regular = 0  they were not: NOTHING was counted — parse them on the host with the exact
 *                 emulation of the reference parser and use katgpu_count_feed.
 * No k-mer spans two calls (records are whole).  On a multi-GPU context the call is collective. */
katgpu_status katgpu_count_feed_fastq(katgpu_table *t, const char *bytes, size_t n, unsigned trim5p,
                                      int *regular);
/* Optional hint: about `expected_kmers` k-mer instances will be fed before
 * katgpu_count_end.  Lets the library size its pending-key buffer once (the
 * partitioned counting path inserts slice by slice; fewer, larger sweeps are
 * cheaper).  Never changes results. */
katgpu_status katgpu_count_reserve(katgpu_table *t, uint64_t expected_kmers);
/* Add `n` (k-mer, count) records to the table: count[keys[i]] += counts[i] — what
 * HashLoader::loadHash does with every record of a jellyfish hash file
 * (`hash->add(reader.key(), reader.val())`, lib/src/jellyfish_helper.cc:176-178;
 * array_base::add, JFI/large_hash_array.hpp:298-302).  keys are 2 bits per base,
 * first base most significant, < 4^k; they are stored as given (a hash dumped
 * from canonical counting already holds canonical k-mers).  HOST pointers. */
katgpu_status katgpu_count_add_pairs(katgpu_table *t, const uint64_t *keys, const uint64_t *counts, size_t n);
/* Finish counting (hash_counter::done, JFI/hash_counter.hpp:173-176).
 * n_distinct / n_total (k-mer instances) may be NULL. */
katgpu_status katgpu_count_end(katgpu_table *t, uint64_t *n_distinct, uint64_t *n_total);
void katgpu_table_free(katgpu_table *t);

/* ---- table walks ------------------------------------------------------------ */

/* Histogram::binSlice + merge (src/histogram.cc:183-199, 146-160): for every
 * distinct k-mer with count v: v < base -> bin 0; v > ceil -> bin nb-1; else
 * bin (v-base)/inc.  bins_out[nb_buckets]. */
katgpu_status katgpu_hist(katgpu_table *t, uint64_t base, uint64_t ceil, uint64_t inc,
                          uint64_t nb_buckets, uint64_t *bins_out);

/* Gcp::analyseSlice + merge (src/gcp.cc:179-197): mx_out[k][cvg_bins+1],
 * row = #G+#C of the stored k-mer, col = min(ceil(count*cvg_scale), cvg_bins).
 * k-mers with GC == k fall outside the k rows and are dropped, as the
 * reference's ThreadedSparseMatrix(k, ...) does (src/gcp.cc:93,
 * lib/include/kat/sparse_matrix.hpp:324-335). */
katgpu_status katgpu_gcp(katgpu_table *t, double cvg_scale, uint32_t cvg_bins, uint64_t *mx_out);

/* Comp::compareSlice + merge (src/comp.cc:387-484) with the CompCounters update
 * rules (lib/src/comp_counters.cc:91-140).  t3 may be NULL (2-input mode); then
 * ends/middle/mixed may be NULL.  Matrices are [d1_bins][d2_bins]; spectra are
 * 4 arrays of min(d1_bins,d2_bins): spectrum1, spectrum2, shared_spectrum1,
 * shared_spectrum2. */
typedef struct katgpu_comp_counters {
    uint64_t hash1_total, hash2_total, hash3_total;
    uint64_t hash1_distinct, hash2_distinct, hash3_distinct;
    uint64_t hash1_only_total, hash2_only_total;
    uint64_t hash1_only_distinct, hash2_only_distinct;
    uint64_t shared_hash1_total, shared_hash2_total, shared_distinct;
} katgpu_comp_counters;

katgpu_status katgpu_comp(katgpu_table *t1, katgpu_table *t2, katgpu_table *t3, double d1_scale,
                          double d2_scale, uint32_t d1_bins, uint32_t d2_bins, uint64_t *main_mx,
                          uint64_t *ends_mx, uint64_t *middle_mx, uint64_t *mixed_mx,
                          katgpu_comp_counters *counters, uint64_t *spectra);

/* JellyfishHelper::getCount (lib/src/jellyfish_helper.cc:189-194): counts_out[i]
 * = count of keys[i] (2 bits per base, first base most significant, as
 * jellyfish::mer_dna packs them), 0 if absent; the probe key is replaced by
 * min(key, revcomp) iff canonicalise. */
katgpu_status katgpu_lookup(katgpu_table *t, const uint64_t *keys, size_t n, int canonicalise,
                            uint64_t *counts_out);

/* Every (key, count) of the table, unordered: what region_slice/eager_slice
 * iteration yields (JFI/large_hash_array.hpp:487-503,
 * JFI/large_hash_iterator.hpp:56-64,194-206).  Writes at most `cap` pairs and
 * returns the total number of distinct k-mers in *n_out. */
katgpu_status katgpu_table_export(katgpu_table *t, uint64_t *keys_out, uint64_t *counts_out,
                                  size_t cap, uint64_t *n_out);

/* The same walk with FilterKmer::inBounds applied on the device (src/filter_kmer.cc:257-309): a k-mer is
 * "in bounds" when low_count <= count <= high_count and low_gc <= #G+#C <= high_gc.  inside != 0
 * returns the k-mers in bounds, inside == 0 those out of bounds (`kat filter kmer`, its --invert
 * and the discard set of --separate). */
katgpu_status katgpu_table_export_filtered(katgpu_table *t, uint64_t low_count, uint64_t high_count, uint32_t low_gc,
                                           uint32_t high_gc, int inside, uint64_t *keys_out, uint64_t *counts_out,
                                           size_t cap, uint64_t *n_out);

/* ---- k-mers of 33..64 bases: two words per key -----------------------------------------
 * jellyfish::mer_dna keeps ceil(2k/64) words per k-mer (JFI/mer_dna.hpp:513-531) and KAT has no limit on k; tables
 * counted with k > 32 ("wide" tables: their own layout and kernels, one GPU) hand keys across the ABI as two arrays:
 * keys_lo = the last 32 bases, keys_hi = the first k-32 bases, 2 bits per base, first base most significant.
 * katgpu_count_begin / _feed* / _end / katgpu_hist / _gcp / _comp / katgpu_table_get_info take wide tables as they are;
 * the three calls that carry keys have _wide forms, which also accept one-word tables (keys_hi is then ignored on
 * input — may be NULL — and zero on output). */
katgpu_status katgpu_lookup_wide(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo, size_t n,
                                 int canonicalise, uint64_t *counts_out);
katgpu_status katgpu_table_export_wide(katgpu_table *t, uint64_t *keys_hi_out, uint64_t *keys_lo_out,
                                       uint64_t *counts_out, size_t cap, uint64_t *n_out);
katgpu_status katgpu_table_export_filtered_wide(katgpu_table *t, uint64_t low_count, uint64_t high_count,
                                                uint32_t low_gc, uint32_t high_gc, int inside, uint64_t *keys_hi_out,
                                                uint64_t *keys_lo_out, uint64_t *counts_out, size_t cap,
                                                uint64_t *n_out);
katgpu_status katgpu_count_add_pairs_wide(katgpu_table *t, const uint64_t *keys_hi, const uint64_t *keys_lo,
                                          const uint64_t *counts, size_t n);

/* ---- table introspection / multi-GPU plumbing ---------------------------------- */

typedef struct katgpu_table_info {
    int k, canonical;
    uint64_t slots;        /* current capacity                                  */
    uint64_t bytes;        /* device bytes held by the table                     */
    uint64_t n_distinct, n_total;
    uint32_t count_bits;   /* width of the in-slot count field                   */
    uint32_t grow_events;  /* number of doublings so far                          */
    uint64_t overflow_keys;/* keys whose count left the in-slot field             */
} katgpu_table_info;
katgpu_status katgpu_table_get_info(katgpu_table *t, katgpu_table_info *out);

/* Testing hook: cap the in-slot count field to `bits` (>=2) for tables created
 * afterwards, to exercise the exact-overflow side table at small sizes.  0 = auto. */
void katgpu_debug_set_count_bits(int bits);

/* ---- multi-GPU jobs ---------------------------------------------
 * One context per GPU; N contexts (N host threads of one process, or N processes —
 * one per GPU, e.g. under torchrun) form a job.  A count histogram is not additive
 * over read shards, so every distinct k-mer is counted by exactly one owner rank:
 * reads are sharded by record, each rank feeds ITS shard, and inside
 * katgpu_count_feed* the extraction kernel stores every k-mer's hashed key straight
 * into the owner's memory over NVLink (peer-mapped receive areas, no NCCL in the data
 * path, no host staging); the owner bins and inserts it with the single-GPU kernels.
 *
 *   katgpu_dist_unique_id   128-byte job id (ncclGetUniqueId); rank 0 makes it, the
 *                           caller hands it to every rank (a broadcast, a file, argv).
 *   katgpu_dist_init        joins the job: NCCL communicator (NCCL is called directly,
 *                           dlopen'ed on first use), receive areas, peer mappings.
 *                           2 <= n_ranks <= 8 (one NVSwitch box).
 *
 * On a job context the calls below are COLLECTIVE — every rank makes them in the same
 * order: katgpu_count_begin, katgpu_count_feed / _feed_device (each rank its own
 * bytes, sizes may differ, 0 is fine), katgpu_count_end (returns the job's totals),
 * katgpu_hist / katgpu_gcp / katgpu_comp / katgpu_lookup (every rank receives the
 * complete result: the local walk is followed by ONE ncclAllReduce(sum, uint64) of
 * the bins / matrix / counters, replacing the reference's merge of per-thread
 * results, src/histogram.cc:146-160, src/gcp.cc:158-177, src/comp.cc:308-326).
 * katgpu_comp needs every input counted canonically (the default); katgpu_table_export
 * returns the rank's own part; katgpu_count_add_pairs is single-GPU only. */
katgpu_status katgpu_dist_unique_id(void *id128);
katgpu_status katgpu_dist_init(katgpu_ctx *ctx, int rank, int n_ranks, const void *id128);
int katgpu_dist_rank(katgpu_ctx *ctx);    /* 0 on a single-GPU context */
int katgpu_dist_n_ranks(katgpu_ctx *ctx); /* 1 on a single-GPU context */
katgpu_status katgpu_dist_barrier(katgpu_ctx *ctx);
/* k-mer starts routed and hashed keys received by this rank since katgpu_dist_init, and
 * route rounds taken part in (8 bytes per key cross NVLink for (N-1)/N of them). */
katgpu_status katgpu_dist_stats(katgpu_ctx *ctx, uint64_t *starts_routed, uint64_t *keys_received,
                                uint64_t *rounds);

/* Measurement aid: the multi-GPU extraction kernel on ONE GPU with n_ranks local receive areas
 * (its cost without NVLink); keys are discarded.  tools/route_probe.py, ncu captures. */
katgpu_status katgpu_debug_route_local(katgpu_table *t, const void *bases_dev, size_t n, int n_ranks,
                                       double *ms, uint64_t *keys);

/* Building blocks of the round-1 exchange (route into caller-provided buckets, insert
 * received keys); kept for callers that move the keys themselves.  Every distinct k-mer must be owned by
 * exactly one rank before histograms become additive:
 *   1. katgpu_route_device: extract + canonicalise + hash this rank's base
 *      stream and write each k-mer's 64-bit hashed key into the bucket of its
 *      owner rank: out_dev[r] (device pointers, possibly peer-mapped memory of
 *      rank r) with capacity cap_keys each; counts_out[r] = keys written.
 *   2. the caller exchanges the buckets (NCCL all-to-all or peer writes);
 *   3. katgpu_insert_hashed_device inserts received hashed keys.
 * After that hist/gcp/comp outputs of the ranks add up (one all-reduce). */
katgpu_status katgpu_route_device(katgpu_table *t, const void *bases_dev, size_t n, int n_ranks,
                                  void *const *out_dev, uint64_t cap_keys, uint64_t *counts_out);
katgpu_status katgpu_insert_hashed_device(katgpu_table *t, const void *hashed_dev, uint64_t n);
/* CUDA IPC export/import of a device buffer (64-byte handle), so that rank r's
 * receive bucket can be written directly over NVLink by its peers. */
katgpu_status katgpu_device_alloc(katgpu_ctx *ctx, size_t bytes, void **out_dev);
void katgpu_device_free(katgpu_ctx *ctx, void *dev);
katgpu_status katgpu_ipc_export(void *dev, void *handle64);
katgpu_status katgpu_ipc_import(const void *handle64, void **out_dev);
void katgpu_ipc_close(void *dev);
katgpu_status katgpu_synchronize(katgpu_ctx *ctx);
/* The library keeps freed device blocks (tables, pending-key buffers) for the next job of the
 * context — the reference likewise keeps its mmap'd arena; this gives them back to the driver. */
katgpu_status katgpu_release_cached_memory(katgpu_ctx *ctx);

/* ---- measurement ------------------------------------------------------------ */

/* Kernel launches issued by this context since creation (bench `gpu_launches`). */
uint64_t katgpu_launch_count(katgpu_ctx *ctx);
/* Per-kernel CUDA-event timing on the library's own stream.  enable=1 starts
 * recording (and clears), katgpu_profile_get returns accumulated milliseconds
 * and launch count for a kernel family: "count", "hist", "gcp", "comp",
 * "route", "insert", "memset". */
katgpu_status katgpu_profile_enable(katgpu_ctx *ctx, int enable);
katgpu_status katgpu_profile_get(katgpu_ctx *ctx, const char *family, double *ms_total,
                                 uint64_t *launches);

/* Deterministic synthetic reads: genome of `genome_len`
 * i.i.d. bases from seed_g; read i is a `read_len` substring at a uniform
 * position, strand uniform, per-base substitution rate sub_rate, per-base N
 * rate n_rate, all drawn from a counter-based RNG keyed by (seed_r, i) so any
 * shard [first_read, first_read+n_reads) is reproducible on its own.  Output
 * is the base stream (read, 'N', read, 'N', ...), n_reads*(read_len+1) bytes,
 * into device memory (katgpu_synth_reads_device) or host memory
 * (katgpu_synth_reads_host: same bytes, computed on the CPU, for the oracle). */
katgpu_status katgpu_synth_reads_device(katgpu_ctx *ctx, uint64_t genome_len, uint64_t seed_g,
                                        uint64_t seed_r, uint64_t first_read, uint64_t n_reads,
                                        uint32_t read_len, double sub_rate, double n_rate,
                                        void *out_dev);
katgpu_status katgpu_synth_reads_host(uint64_t genome_len, uint64_t seed_g, uint64_t seed_r,
                                      uint64_t first_read, uint64_t n_reads, uint32_t read_len,
                                      double sub_rate, double n_rate, char *out);
katgpu_status katgpu_memcpy_d2h(katgpu_ctx *ctx, void *dst_host, const void *src_dev, size_t n);
katgpu_status katgpu_memcpy_h2d(katgpu_ctx *ctx, void *dst_dev, const void *src_host, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* KATGPU_H */

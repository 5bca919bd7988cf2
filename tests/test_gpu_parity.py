"""GPU parity: libkatgpu (through the C-ABI) vs the CPU oracle on identical
seeded inputs.  Integer work: the bar is bit-exact equality."""
import numpy as np
import pytest

import oracle as O
import seqgen
import kat_b200
from kat_b200 import binding as B

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["direct", "partition", "tiled"])
def count_path(request, monkeypatch):
    """Every case runs through the three counting paths of libkatgpu: extract+insert in
    read order ("direct"), partition-by-slice + L2-resident sweep ("partition"), and two binning
    passes + shared-memory table tiles ("tiled"; tables too small for it fall back to "direct")."""
    monkeypatch.setenv("KATGPU_COUNT_PATH", request.param)
    return request.param


def ora_table(chunks, k, canonical=True):
    return O.Table.from_chunks([bytes(c) for c in chunks], k, canonical)


def gpu_table(ctx, chunks, k, canonical=True, hash_size=1 << 16, allow_grow=True):
    return ctx.count([np.ascontiguousarray(c) for c in chunks], k, canonical, hash_size, allow_grow)


def assert_same_multiset(gt, ot):
    keys, counts, n = gt.export()
    okeys, ocounts = ot.export()
    assert n == ot.distinct == gt.n_distinct
    assert gt.n_total == ot.total
    assert np.array_equal(keys, okeys)
    assert np.array_equal(counts, ocounts)


@pytest.mark.parametrize("k", [1, 2, 5, 13, 15, 16, 17, 21, 27, 31, 32])
@pytest.mark.parametrize("canonical", [True, False])
def test_count_matches_oracle(ctx, k, canonical):
    s = seqgen.random_stream(60_000, seed=100 + k)
    gt = gpu_table(ctx, [s], k, canonical)
    ot = ora_table([s], k, canonical)
    assert_same_multiset(gt, ot)
    base, bins = gt.hist()
    obase, obins = ot.hist()
    assert base == obase and np.array_equal(bins, obins)
    assert int(bins.sum()) == ot.distinct


@pytest.mark.parametrize("n", [0, 1, 26, 27, 28, 31, 32, 33, 8191, 8192, 8193, 8192 + 31, 8192 + 32, 8192 + 33,
                               3 * 8192, 3 * 8192 + 17])
def test_ragged_sizes(ctx, n):
    """Empty, shorter than k, and sizes around the 8 KiB TMA tile / 32 B halo."""
    s = seqgen.random_stream(n, seed=n + 7, n_rate=0.002)
    gt = gpu_table(ctx, [s], 27)
    ot = ora_table([s], 27)
    assert_same_multiset(gt, ot)


def test_no_kmer_spans_feeds(ctx):
    a = seqgen.random_stream(5000, 1, n_rate=0.0, junk_rate=0.0)
    b = seqgen.random_stream(7000, 2, n_rate=0.0, junk_rate=0.0)
    gt = gpu_table(ctx, [a, b], 21)
    assert_same_multiset(gt, ora_table([a, b], 21))
    joined = gpu_table(ctx, [np.concatenate([a, b])], 21)
    assert joined.n_total == gt.n_total + 20


def test_device_feed_equals_host_feed(ctx):
    s = seqgen.genome_reads(50_000, 4000, 150, seed=5)
    d = ctx.device_alloc(s.size)
    try:
        ctx.h2d(d, s)
        t = ctx.count_begin(27, True, 1 << 18)
        t.feed_device(d, s.size).end()
        assert_same_multiset(t, ora_table([s], 27))
        # unaligned device pointer: falls back to the guarded loader
        t2 = ctx.count_begin(27, True, 1 << 18)
        t2.feed_device(d + 3, s.size - 3).end()
        assert_same_multiset(t2, ora_table([s[3:]], 27))
    finally:
        ctx.device_free(d)


def test_high_multiplicity_and_hist_params(ctx):
    s = np.concatenate([seqgen.genome_reads(20_000, 6000, 150, seed=9), seqgen.low_complexity(400, 150, seed=3)])
    gt = gpu_table(ctx, [s], 21)
    ot = ora_table([s], 21)
    assert_same_multiset(gt, ot)
    for low, high, inc in [(1, 10000, 1), (2, 6, 2), (5, 50, 5), (1, 3, 1), (10, 10, 1), (3, 100000, 7)]:
        base, bins = gt.hist(low, high, inc)
        obase, obins = ot.hist(low, high, inc)
        assert base == obase and np.array_equal(bins, obins), (low, high, inc)


def test_count_overflow_side_table(ctx):
    """Force a 3-bit in-slot count field: every 8th increment carries into the
    exact side table (the reference's large-value chaining, large_hash_array.hpp:668-686)."""
    B.lib().katgpu_debug_set_count_bits(3)
    try:
        s = np.concatenate([seqgen.low_complexity(300, 104, seed=1), seqgen.genome_reads(3000, 2000, 100, seed=2)])
        gt = gpu_table(ctx, [s], 5)
        ot = ora_table([s], 5)
        assert gt.info()["count_bits"] == 3 and gt.info()["overflow_keys"] > 0
        assert_same_multiset(gt, ot)
        assert np.array_equal(gt.hist()[1], ot.hist()[1])
        assert np.array_equal(gt.gcp(), ot.gcp())
        g2 = gpu_table(ctx, [s], 17)
        o2 = ora_table([s], 17)
        assert_same_multiset(g2, o2)
        # growth with overflowed keys present
        g3 = gpu_table(ctx, [s], 17, hash_size=4096)
        assert g3.info()["grow_events"] > 0
        assert_same_multiset(g3, o2)
    finally:
        B.lib().katgpu_debug_set_count_bits(0)


def test_exact_large_count(ctx):
    """300 reads of 104 x 'A', k=5: one k-mer with count 30000."""
    s = np.tile(np.frombuffer(b"A" * 104 + b"N", dtype=np.uint8), 300)
    gt = gpu_table(ctx, [s], 5)
    keys, counts, n = gt.export()
    assert n == 1 and keys[0] == 0 and counts[0] == 30000


def test_table_growth_and_hash_full(ctx, count_path):
    s = seqgen.random_stream(400_000, seed=77, n_rate=0.001)
    ot = ora_table([s], 27)
    gt = gpu_table(ctx, [s], 27, hash_size=4096, allow_grow=True)
    assert gt.info()["grow_events"] >= (3 if count_path == "direct" else 1)  # the partitioned path sizes once from its HLL estimate
    assert_same_multiset(gt, ot)
    with pytest.raises(kat_b200.KatGpuError) as e:
        gpu_table(ctx, [s], 27, hash_size=4096, allow_grow=False)
    assert e.value.status == -3 and "Hash full" in str(e.value)


@pytest.mark.parametrize("k,scale,bins", [(17, 1.0, 1000), (5, 0.5, 3), (21, 0.1, 50), (27, 2.5, 1000), (9, 1.0, 20)])
def test_gcp(ctx, k, scale, bins):
    s = np.concatenate([seqgen.genome_reads(30_000, 5000, 150, seed=k), seqgen.low_complexity(50, 150, seed=k),
                        np.frombuffer(b"N" + b"GC" * 40 + b"N" + b"G" * 60 + b"N", dtype=np.uint8)])
    gt = gpu_table(ctx, [s], k)
    ot = ora_table([s], k)
    assert np.array_equal(gt.gcp(scale, bins), ot.gcp(scale, bins))


@pytest.mark.parametrize("canon", [(True, True), (False, False), (False, True), (True, False)])
@pytest.mark.parametrize("scales", [(1.0, 1.0, 1001, 1001), (0.5, 2.0, 40, 60), (1.0, 0.25, 300, 20)])
def test_comp_two_inputs(ctx, canon, scales):
    d1s, d2s, d1b, d2b = scales
    a = seqgen.genome_reads(40_000, 6000, 150, seed=11)
    b = np.concatenate([seqgen.genome_reads(40_000, 3000, 150, seed=11)[: 3000 * 151 // 2],
                        seqgen.genome_reads(30_000, 2000, 150, seed=12)])
    k = 21
    g1, g2 = gpu_table(ctx, [a], k, canon[0]), gpu_table(ctx, [b], k, canon[1])
    o1, o2 = ora_table([a], k, canon[0]), ora_table([b], k, canon[1])
    gr = B.comp(g1, g2, None, d1s, d2s, d1b, d2b)
    orr = O.comp(o1, o2, None, d1s, d2s, d1b, d2b)
    assert gr["counters"] == orr["counters"]
    assert np.array_equal(gr["spectra"], orr["spectra"])
    assert np.array_equal(gr["main"], orr["main"])


def test_comp_three_inputs(ctx):
    k = 17
    a = seqgen.genome_reads(30_000, 5000, 150, seed=21)
    b = seqgen.genome_reads(30_000, 400, 150, seed=21, sub_rate=0.0)
    c = seqgen.genome_reads(30_000, 300, 150, seed=22, sub_rate=0.0)
    g = [gpu_table(ctx, [x], k) for x in (a, b, c)]
    o = [ora_table([x], k) for x in (a, b, c)]
    gr = B.comp(*g, 1.0, 1.0, 200, 150)
    orr = O.comp(*o, 1.0, 1.0, 200, 150)
    assert gr["counters"] == orr["counters"]
    for name in ("main", "ends", "middle", "mixed", "spectra"):
        assert np.array_equal(gr[name], orr[name]), name


def test_lookup(ctx):
    s = seqgen.genome_reads(20_000, 3000, 150, seed=31)
    for canonical in (True, False):
        gt = gpu_table(ctx, [s], 27, canonical)
        ot = ora_table([s], 27, canonical)
        keys, _ = ot.export()
        rng = np.random.default_rng(1)
        probe = np.concatenate([keys[:2000], rng.integers(0, 1 << 54, 2000, dtype=np.uint64)])
        for canon_probe in (False, True):
            got = gt.lookup(probe, canon_probe)
            want = np.array([ot.lookup(int(x), canon_probe) for x in probe], dtype=np.uint64)
            assert np.array_equal(got, want)


@pytest.mark.parametrize("k", [5, 21, 32])
def test_export_filtered(ctx, k):
    """katgpu_table_export_filtered = FilterKmer::inBounds on the device (src/filter_kmer.cc:296-309): the in-bounds and
    out-of-bounds sets partition the table and equal the oracle's multiset filtered on the host; counts beyond the
    in-slot field (overflow side table) are filtered by their full value."""
    s = np.concatenate([seqgen.genome_reads(5_000, 4000, 150, seed=77), np.frombuffer(b"N" + b"A" * 700 + b"N" + b"GC" * 300, dtype=np.uint8)])
    B.lib().katgpu_debug_set_count_bits(4)
    try:
        gt = gpu_table(ctx, [s], k, True)
    finally:
        B.lib().katgpu_debug_set_count_bits(0)
    keys, counts = ora_table([s], k, True).export()
    gc = np.zeros(keys.size, dtype=np.uint32)
    for b in range(k):
        c = (keys >> np.uint64(2 * b)) & np.uint64(3)
        gc += ((c == 1) | (c == 2)).astype(np.uint32)
    for lo_c, hi_c, lo_g, hi_g in ((1, 10000, 0, 100), (2, 5, k // 3, 2 * k // 3), (16, 1 << 40, 0, k), (3, 3, k, k)):
        inb = (counts >= lo_c) & (counts <= hi_c) & (gc >= lo_g) & (gc <= hi_g)
        for inside, sel in ((True, inb), (False, ~inb)):
            gk, gc_ = gt.export_filtered(lo_c, hi_c, lo_g, hi_g, inside)
            assert np.array_equal(gk, keys[sel]) and np.array_equal(gc_, counts[sel]), (k, lo_c, hi_c, lo_g, hi_g, inside)


def test_synth_reads_device_equals_host(ctx):
    n_reads, L = 5000, 150
    host = B.synth_reads_host(1_000_000, 1, 2, 12345, n_reads, L)
    d = ctx.device_alloc(host.size)
    try:
        ctx.synth_reads_device(d, 1_000_000, 1, 2, 12345, n_reads, L)
        dev = np.empty_like(host)
        ctx.d2h(dev, d)
        assert np.array_equal(host, dev)
        t = ctx.count_begin(27, True, 1 << 20)
        t.feed_device(d, host.size).end()
        assert_same_multiset(t, ora_table([host], 27))
    finally:
        ctx.device_free(d)


@pytest.mark.parametrize("R", [2, 3, 4, 8])
def test_route_then_insert_equals_direct(ctx, R):
    """Single-process model of the multi-GPU exchange: route to R owner buckets,
    insert every bucket into its own table; the union must equal direct counting
    and every key must land in exactly one table."""
    s = seqgen.genome_reads(60_000, 8000, 150, seed=41)
    k = 27
    d = ctx.device_alloc(s.size)
    cap = s.size  # generous
    outs = [ctx.device_alloc(cap * 8) for _ in range(R)]
    try:
        ctx.h2d(d, s)
        router = ctx.count_begin(k, True, 1 << 12)
        counts = router.route_device(d, s.size, outs, cap)
        ot = ora_table([s], k)
        assert sum(counts) == ot.total
        tables = []
        for r in range(R):
            t = ctx.count_begin(k, True, 1 << 18)
            t.insert_hashed_device(outs[r], counts[r])
            t.end()
            tables.append(t)
        allk, allc = [], []
        for t in tables:
            kk, cc, n = t.export()
            allk.append(kk); allc.append(cc)
        kk = np.concatenate(allk); cc = np.concatenate(allc)
        assert len(np.unique(kk)) == len(kk) == ot.distinct  # one owner per key
        order = np.argsort(kk)
        okeys, ocounts = ot.export()
        assert np.array_equal(kk[order], okeys) and np.array_equal(cc[order], ocounts)
        # additivity of the reductions after ownership partitioning
        bins = sum(t.hist()[1] for t in tables)
        assert np.array_equal(bins, ot.hist()[1])
        assert np.array_equal(sum(t.gcp() for t in tables), ot.gcp())
    finally:
        ctx.device_free(d)
        for o in outs:
            ctx.device_free(o)


@pytest.mark.parametrize("k", [5, 21, 32])
def test_add_pairs_is_hash_add(ctx, k):
    """katgpu_count_add_pairs == `hash->add(key, val)` per record (HashLoader::loadHash): loading the
    records of a counted table into a fresh one — in two overlapping halves, so counts add — gives
    the oracle's counts; counts beyond the in-slot field stay exact (forced 3-bit field)."""
    reads = seqgen.genome_reads(3000, 1500, 90, seed=k)
    ot = ora_table([reads], k)
    okeys, ocounts = ot.export()
    B.lib().katgpu_debug_set_count_bits(3)
    try:
        t = ctx.count_begin(k, True, 1 << 10, True)
        half = len(okeys) // 2
        t.add_pairs(okeys[:half], ocounts[:half])
        t.add_pairs(okeys[half:], ocounts[half:] + np.uint64(1 << 34))  # large counts: carried exactly
        t.add_pairs(okeys[: half // 2], np.full(half // 2, 3, dtype=np.uint64))  # a second add of the same keys
        t.add_pairs(np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64))
        t.end()
    finally:
        B.lib().katgpu_debug_set_count_bits(0)
    want = ocounts.copy()
    want[half:] += np.uint64(1 << 34)
    want[: half // 2] += np.uint64(3)
    keys, counts, n = t.export()
    assert n == len(okeys) and np.array_equal(keys, okeys) and np.array_equal(counts, want)
    assert t.n_total == int(want.sum())
    _, bins = t.hist(1, 10000, 1)
    assert int(bins.sum()) == n


@pytest.mark.parametrize("slots_log2,k", [(22, 27), (26, 21), (26, 32), (24, 15)])
def test_larger_tables_many_tiles(ctx, slots_log2, k):
    """Tables of 2^22..2^26 slots = 2^9..2^13 shared-memory tiles: both binning levels have real
    fan-out, feeds arrive in several launches (unaligned segment tails), and a low-complexity
    tail makes full home buckets, count-field carries (forced 6-bit field) and the side list real."""
    s1 = seqgen.genome_reads(120_000, 9000, 150, seed=slots_log2 + k)
    s2 = seqgen.random_stream(700_001, seed=k, n_rate=0.002)
    s3 = np.concatenate([seqgen.low_complexity(400, 150, seed=3), seqgen.genome_reads(120_000, 5000, 150, seed=slots_log2 + k)])
    B.lib().katgpu_debug_set_count_bits(6)
    try:
        gt = gpu_table(ctx, [s1, s2, s3], k, hash_size=1 << slots_log2, allow_grow=True)
    finally:
        B.lib().katgpu_debug_set_count_bits(0)
    assert_same_multiset(gt, ora_table([s1, s2, s3], k))


@pytest.mark.parametrize("env", [{"KATGPU_TILED_GROUP_MB": "1"},      # many small groups of slices
                                 {"KATGPU_TILED_GENERAL": "1"},        # level-1 segments inserted the general way
                                 {"KATGPU_SIDE_CAP": "64"},            # side list too small: tiles deferred, rings spill to direct inserts
                                 {"KATGPU_SIDE_CAP": "20000", "KATGPU_TILED_GROUP_MB": "2"}],
                         ids=["small_groups", "general", "tiny_side_list", "small_side_list_and_groups"])
def test_tiled_rare_paths(ctx, env, count_path, monkeypatch):
    """The slow-but-exact exits of the tiled path (tile_kernels.cu): group-wise level-2 buffers,
    the general fallback, deferred tiles and a full side list must all give the oracle's multiset."""
    if count_path != "tiled":
        pytest.skip("tiled path only")
    for k_, v in env.items():
        monkeypatch.setenv(k_, v)
    k = 21
    s1 = seqgen.genome_reads(150_000, 12000, 150, seed=5)
    s2 = np.concatenate([seqgen.low_complexity(300, 150, seed=9), seqgen.random_stream(300_000, seed=2)])
    B.lib().katgpu_debug_set_count_bits(5)
    try:
        gt = gpu_table(ctx, [s1, s2], k, hash_size=1 << 22, allow_grow=True)
    finally:
        B.lib().katgpu_debug_set_count_bits(0)
    assert_same_multiset(gt, ora_table([s1, s2], k))


def test_reference_hash_fixture_queries(ctx):
    """tests/check_jellyfish.cc:62-116 on the GPU: load the reference's own hash fixture
    (ecoli.header.jf27: 1889 records) record by record — HashLoader::loadHash — and query the four
    named 27-mers: 3/1/1/1 as stored, 3/1/0/0 with the probe canonicalised (JellyfishHelper::getCount)."""
    import gzip
    import os
    import golden_util as G
    recs = [l.split() for l in gzip.open(os.path.join(G.GOLDEN, "jf", "ref_ecoli_header_k27.dump.txt.gz"), "rt").read().splitlines()]
    assert len(recs) == 1889
    keys = np.array([seqgen.encode_kmer(a) for a, _ in recs], dtype=np.uint64)
    counts = np.array([int(c) for _, c in recs], dtype=np.uint64)
    t = ctx.count_begin(27, False, 2 * len(recs), True)
    t.add_pairs(keys, counts).end()
    assert t.n_distinct == 1889 and t.n_total == int(counts.sum())
    named = ["AGCTTTTCATTCTGACTGCAACGGGCA", "GCATAGCGCACAGACAGATAAAAATTA", "AATGAAAAAGGCGAACTGGTGGTGCTT", "CTCACCAATGTACATGGCCTTAATCTG"]
    probe = np.array([seqgen.encode_kmer(s) for s in named], dtype=np.uint64)
    assert list(t.lookup(probe, False)) == [3, 1, 1, 1]
    assert list(t.lookup(probe, True)) == [3, 1, 0, 0]


def test_c1_named_kmer_on_gpu(ctx):
    """tests/check_jellyfish.cc:118-135 through the C-ABI: count EcoliK12 (committed fixture) at
    k=27 canonical, query AGCTTTTCATTCTGACTGCAACGGGCA -> 1; totals 4,641,626 / 4,550,880."""
    import os
    import golden_util as G
    chunks = O.parse_files([os.path.join(G.GOLDEN, "c1", "EcoliK12.fasta.gz")], 27)
    t = ctx.count([np.frombuffer(c, dtype=np.uint8) for c in chunks], 27, True, 1 << 24)
    assert (t.n_total, t.n_distinct) == (4641626, 4550880)
    assert list(t.lookup(np.array([seqgen.encode_kmer("AGCTTTTCATTCTGACTGCAACGGGCA")], dtype=np.uint64), True)) == [1]


def _multiset_checksums(keys, counts):
    """Order-independent fingerprints of a (key, count) multiset (uint64 arithmetic wraps)."""
    keys, counts = keys.astype(np.uint64), counts.astype(np.uint64)
    m = np.uint64(0x9E3779B97F4A7C15)
    mixed = (keys ^ (keys >> np.uint64(29))) * m
    return (int(len(keys)), int(counts.sum(dtype=np.uint64)), int((mixed * counts).sum(dtype=np.uint64)),
            int(np.bitwise_xor.reduce(mixed + counts * m)) if len(keys) else 0)


@pytest.mark.parametrize("slots_log2", [32, 30])
def test_tiled_path_at_benchmark_geometry(ctx, slots_log2, count_path):
    """The benchmarked path at the benchmark's table geometry (bench.py: 2^32 slots -> 512 level-1
    slices, 1024 tiles per slice, several level-2 groups; 2^30 for the 2-GPU setting): 2 M synthetic
    150 bp reads generated ON THE DEVICE (the bench's generator, ~250 M 27-mers), counted through the
    tiled kernels, against the oracle on the same bytes: totals, hist, gcp, order-independent
    checksums of the full exported multiset, and 200k point lookups."""
    if count_path != "tiled":
        pytest.skip("the benchmarked (tiled) path only")
    n_reads, L, k = 2_000_000, 150, 27
    n_bytes = n_reads * (L + 1)
    d = ctx.device_alloc(n_bytes + 64)
    try:
        ctx.synth_reads_device(d, 200_000_000, 1, 2, 0, n_reads, L, 0.005, 0.0001)
        host = np.empty(n_bytes, dtype=np.uint8)
        ctx.d2h(host, d)
        t = ctx.count_begin(k, True, 1 << slots_log2, True)
        t.feed_device(d, n_bytes).end()
    finally:
        ctx.device_free(d)
    assert np.array_equal(host[: 3 * (L + 1)], O.synth_reads(200_000_000, 1, 2, 0, 3, L))  # same generator as the reference arm's
    ot = O.Table.from_chunks([host], k, True)
    assert (t.n_total, t.n_distinct) == (ot.total, ot.distinct)
    assert t.info()["slots"] == 1 << slots_log2
    assert np.array_equal(t.hist()[1], ot.hist()[1])
    assert np.array_equal(t.gcp(), ot.gcp())
    okeys, ocounts = ot.export()
    keys, counts, n = t.export(sort=False)
    assert n == len(okeys)
    assert _multiset_checksums(keys, counts) == _multiset_checksums(okeys, ocounts)
    rng = np.random.default_rng(7)
    pick = rng.integers(0, len(okeys), 200_000)
    probe = np.concatenate([okeys[pick], rng.integers(0, 1 << 54, 1000, dtype=np.uint64)])
    want = np.concatenate([ocounts[pick], np.array([ot.lookup(int(x)) for x in probe[-1000:]], dtype=np.uint64)])
    assert np.array_equal(t.lookup(probe, False), want)
    t.free()


# ---- FASTQ records parsed on the device (katgpu_count_feed_fastq, parse_kernels.cu) ---------------
def _regular_fastq(rng, n_records, crlf=False, min_len=1):
    """4-line FASTQ in the regular layout, with everything that layout allows: ragged read lengths
    (1..300), long and empty-ish headers, '@' and '+' opening quality lines, 'N'/IUPAC/lower-case
    bases, repeated header on the '+' line."""
    out = bytearray()
    alpha = np.frombuffer(b"ACGTACGTACGTNacgtRY", dtype=np.uint8)
    qual = np.frombuffer(b"@+I#5?!~", dtype=np.uint8)
    for i in range(n_records):
        L = int(rng.integers(min_len, 301)) if i % 7 else int(rng.integers(min_len, min_len + 5))
        seq = alpha[rng.integers(0, len(alpha), L)].tobytes()
        if seq[:1] in (b"@", b"+"):
            seq = b"A" + seq[1:]
        hdr = b"@r%d %s" % (i, b"x" * int(rng.integers(0, 70))) if i % 11 else b"@"
        q = qual[rng.integers(0, len(qual), L)].tobytes()
        plus = b"+" + (hdr[1:] if i % 3 == 0 else b"")
        eol = b"\r\n" if crlf else b"\n"
        if crlf:  # '\r' is part of the line: the quality line must match the sequence line's length including it
            out += hdr + eol + seq + eol + plus + eol + q + eol
        else:
            out += hdr + eol + seq + eol + plus + eol + q + eol
    return bytes(out)


@pytest.mark.parametrize("k,trim,crlf", [(27, 0, False), (5, 0, False), (21, 3, False), (31, 0, True), (17, 7, True)])
def test_feed_fastq_device_parse_matches_reference_parser(ctx, k, trim, crlf, count_path):
    """Raw 4-line FASTQ bytes parsed on the GPU give exactly the k-mers of the oracle's emulation of
    the reference parser (JFI/mer_overlap_sequence_parser.hpp:218-289), with and without 5' trim."""
    rng = np.random.default_rng(100 * k + trim)
    # (a read no longer than the trim is outside the regular layout: the reference's ignore(trim) would eat its newline)
    raw = _regular_fastq(rng, 20000, crlf, trim + 1 + (1 if crlf else 0))
    stream, st = O.parse_file(raw, k, trim)
    assert st == 0
    t = ctx.count_begin(k, True, 1 << 22, True)
    assert t.feed_fastq(raw, trim) is True
    # second call: another block of whole records (no k-mer spans calls), different size -> buffers regrow
    raw2 = _regular_fastq(rng, 45000, crlf, trim + 1 + (1 if crlf else 0))
    assert t.feed_fastq(raw2, trim) is True
    t.end()
    stream2, st2 = O.parse_file(raw2, k, trim)
    assert st2 == 0
    assert_same_multiset(t, ora_table([stream, stream2], k))


@pytest.mark.parametrize("what", ["multi_line", "no_final_newline", "length_mismatch", "plus_in_sequence", "missing_plus",
                                  "three_lines", "empty_sequence", "trim_too_long", "fasta", "sequence_starts_with_at"])
def test_feed_fastq_refuses_irregular_input_and_counts_nothing(ctx, what):
    """Every precondition of the regular layout is checked on the device; a violation anywhere means
    regular = 0 and an untouched table, and the host route (exact emulation) then gives the result."""
    good = b"@a\nACGTACGTAC\n+\nIIIIIIIIII\n" * 50
    bad = {"multi_line": b"@r\nACGTA\nCGTAC\n+\nIIIII\nIIIII\n", "no_final_newline": b"@r\nACGTACGTAC\n+\nIIIIIIIIII",
           "length_mismatch": b"@r\nACGTACGTAC\n+\nIIIIIIIII\n", "plus_in_sequence": b"@r\nACGT+CGTAC\n+\nIIIIIIIIII\n",
           "missing_plus": b"@r\nACGTACGTAC\n-\nIIIIIIIIII\n", "three_lines": b"@r\nACGTACGTAC\n+\n",
           "empty_sequence": b"@r\n\n+\n\n", "trim_too_long": b"@r\nACG\n+\nIII\n", "fasta": b">r\nACGTACGTAC\n>s\nACGTACGTAC\n",
           "sequence_starts_with_at": b"@r\n@CGTACGTAC\n+\nIIIIIIIIII\n"}[what]
    raw = good + bad if what != "fasta" else bad * 2
    trim = 3 if what == "trim_too_long" else 0
    t = ctx.count_begin(5, True, 1 << 16, True)
    assert t.feed_fastq(raw, trim) is False
    assert t.info()["n_total"] == 0 and t.info()["n_distinct"] == 0
    stream, st = O.parse_file(raw, 5, trim)
    if st != -1:
        t.feed(stream)
    t.end()
    assert_same_multiset(t, ora_table([stream] if st != -1 else [], 5))
    assert ctx.count_begin(5, True, 1 << 16, True).feed_fastq(b"") is True  # nothing to do is fine

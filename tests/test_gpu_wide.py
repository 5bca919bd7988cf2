"""GPU parity for mer lengths 33..64 (two-word keys: kat_b200/csrc/wide.cuh, wide_kernels.cu, katgpu_wide.cu) against
the CPU oracle, which is pinned to the reference at these k by tests/test_oracle_pinning.py
(test_oracle_reproduces_reference_output, the k = 33..64 cases).  Integer work: bit-exact.  The same device code runs on the CPU in
tests/test_wide_host_emulation.py."""
import numpy as np
import pytest

import oracle as O
import seqgen
import kat_b200
from kat_b200 import binding as B

pytestmark = pytest.mark.gpu

KS = [33, 40, 48, 63, 64]


def ora_table(chunks, k, canonical=True):
    return O.Table.from_chunks([bytes(c) for c in chunks], k, canonical)


def gpu_table(ctx, chunks, k, canonical=True, hash_size=1 << 16, allow_grow=True):
    return ctx.count([np.ascontiguousarray(c) for c in chunks], k, canonical, hash_size, allow_grow)


def assert_same_multiset(gt, ot):
    hi, lo, cn, n = gt.export_wide()
    ohi, olo, ocn = ot.export() if ot.k > 32 else (np.zeros(ot.distinct, dtype=np.uint64),) + ot.export()
    assert n == ot.distinct == gt.n_distinct
    assert gt.n_total == ot.total
    assert np.array_equal(hi, ohi) and np.array_equal(lo, olo) and np.array_equal(cn, ocn)


def mixed_stream(seed):
    """reads with errors, Ns, junk bytes and lower case; long poly-T / poly-A runs (T x 64 is the key that equals the
    table's empty marker); GC-only repeats (gcp drops GC == k); records shorter than k"""
    tail = b"N" + b"T" * 300 + b"N" + b"a" * 200 + b"N" + b"acgtACGTRYacgt" * 20 + b"N" + b"GC" * 90 + b"N" + b"ACGT" * 9 + b"N"
    return np.concatenate([seqgen.genome_reads(20_000, 3000, 150, seed=seed), seqgen.random_stream(30_000, seed=seed + 1),
                           np.frombuffer(tail, dtype=np.uint8)])


@pytest.mark.parametrize("k", KS)
@pytest.mark.parametrize("canonical", [True, False])
def test_wide_count_hist_gcp_match_oracle(ctx, k, canonical):
    s = mixed_stream(200 + k)
    gt = gpu_table(ctx, [s], k, canonical)
    ot = ora_table([s], k, canonical)
    assert_same_multiset(gt, ot)
    base, bins = gt.hist()
    obase, obins = ot.hist()
    assert base == obase and np.array_equal(bins, obins) and int(bins.sum()) == ot.distinct
    assert np.array_equal(gt.hist(2, 6, 2)[1], ot.hist(2, 6, 2)[1])
    for scale, cbins in ((1.0, 1000), (0.5, 7)):
        assert np.array_equal(gt.gcp(scale, cbins), ot.gcp(scale, cbins))
    info = gt.info()
    assert (info["k"], info["canonical"], info["n_distinct"], info["n_total"]) == (k, int(canonical), ot.distinct, ot.total)
    if k == 64 and not canonical:
        assert ot.lookup((1 << 128) - 1) == 300 - 64 + 1
        assert int(gt.lookup_wide([2**64 - 1], [2**64 - 1])[0]) == 300 - 64 + 1


@pytest.mark.parametrize("n", [0, 1, 32, 33, 34, 63, 64, 65, 9215, 9216, 9217, 9216 + 63, 9216 + 64, 9216 + 65, 3 * 9216, 3 * 9216 + 17])
def test_wide_ragged_sizes(ctx, n):
    """Empty, shorter than k, and sizes around the 9216-start TMA tile and its 64-byte halo."""
    s = seqgen.random_stream(n, seed=n + 7, n_rate=0.002)
    for k in (33, 64):
        assert_same_multiset(gpu_table(ctx, [s], k), ora_table([s], k))


def test_wide_multi_feed_device_and_unaligned(ctx):
    """Several feeds (no k-mer spans two), a device-resident stream (TMA tiles) and an unaligned one (guarded loads)."""
    k = 41
    chunks = [seqgen.genome_reads(10_000, 700, 150, seed=s) for s in (1, 2, 3)] + [seqgen.random_stream(50, seed=9), np.zeros(0, dtype=np.uint8)]
    assert_same_multiset(gpu_table(ctx, chunks, k), ora_table(chunks, k))
    s = chunks[0]
    d = ctx.device_alloc(s.size + 64)
    try:
        ctx.h2d(d, s)
        for off in (0, 1, 16, 37):
            t = ctx.count_begin(k, True, 1 << 16)
            t.feed_device(d + off, s.size - off).end()
            assert_same_multiset(t, ora_table([s[off:]], k))
    finally:
        ctx.device_free(d)


def test_wide_growth_and_hash_full(ctx):
    s = seqgen.random_stream(400_000, seed=77, n_rate=0.001)
    ot = ora_table([s], 47)
    gt = gpu_table(ctx, [s], 47, hash_size=4096, allow_grow=True)
    assert gt.info()["grow_events"] >= 3
    assert_same_multiset(gt, ot)
    with pytest.raises(kat_b200.KatGpuError) as e:
        gpu_table(ctx, [s], 47, hash_size=4096, allow_grow=False)
    assert e.value.status == -3 and "Hash full" in str(e.value)


def test_wide_larger_job_device_synth(ctx):
    """2 M reads of 150 bases at k = 41 (220 M k-mer instances) generated on the device, fed in one call: totals, histogram
    and GC matrix against the oracle on an eighth of it (the oracle sorts 16-byte keys), plus the full job's invariants."""
    k, n_reads, L = 41, 2_000_000, 150
    n = n_reads * (L + 1)
    d = ctx.device_alloc(n + 64)
    try:
        ctx.synth_reads_device(d, 2_000_000, 1, 2, 0, n_reads, L)
        t = ctx.count_begin(k, True, 1 << 27)
        t.feed_device(d, n).end()
        base, bins = t.hist()
        lab = base + np.arange(bins.size, dtype=np.uint64)
        assert int(bins.sum()) == t.n_distinct and int((lab[:-1] * bins[:-1]).sum()) <= t.n_total
        if int(bins[-1]) == 0:
            assert int((lab * bins).sum()) == t.n_total
        q = n_reads // 8
        host = B.synth_reads_host(2_000_000, 1, 2, 0, q, L)
        t4 = ctx.count_begin(k, True, 1 << 20)  # has to grow on the way
        t4.feed_device(d, q * (L + 1)).end()
        ot = ora_table([host], k)
        assert (t4.n_distinct, t4.n_total) == (ot.distinct, ot.total)
        assert np.array_equal(t4.hist()[1], ot.hist()[1]) and np.array_equal(t4.gcp(), ot.gcp())
        t.free()
        t4.free()
    finally:
        ctx.device_free(d)


@pytest.mark.parametrize("canon", [(True, True), (False, False), (False, True), (True, False)])
def test_wide_comp_two_inputs(ctx, canon):
    k = 35
    a = np.concatenate([seqgen.genome_reads(30_000, 4000, 150, seed=11), seqgen.low_complexity(40, 150, seed=3)])
    b = seqgen.genome_reads(30_000, 2500, 150, seed=11, sub_rate=0.02)
    ga, gb = gpu_table(ctx, [a], k, canon[0]), gpu_table(ctx, [b], k, canon[1])
    oa, ob = ora_table([a], k, canon[0]), ora_table([b], k, canon[1])
    for scales in ((1.0, 1.0, 1001, 1001), (0.5, 2.0, 40, 60)):
        gr, orr = B.comp(ga, gb, None, *scales), O.comp(oa, ob, None, *scales)
        assert gr["counters"] == orr["counters"]
        for name in ("main", "spectra"):
            assert np.array_equal(gr[name], orr[name]), name


def test_wide_comp_three_inputs(ctx):
    k = 50
    a = seqgen.genome_reads(30_000, 5000, 150, seed=21)
    b = seqgen.genome_reads(30_000, 400, 150, seed=21, sub_rate=0.0)
    c = seqgen.genome_reads(30_000, 300, 150, seed=22, sub_rate=0.0)
    g = [gpu_table(ctx, [x], k) for x in (a, b, c)]
    o = [ora_table([x], k) for x in (a, b, c)]
    gr, orr = B.comp(*g, 1.0, 1.0, 200, 150), O.comp(*o, 1.0, 1.0, 200, 150)
    assert gr["counters"] == orr["counters"]
    for name in ("main", "ends", "middle", "mixed", "spectra"):
        assert np.array_equal(gr[name], orr[name]), name


@pytest.mark.parametrize("k", [33, 47, 64])
def test_wide_lookup(ctx, k):
    s = mixed_stream(31)
    rng = np.random.default_rng(k)
    mask = (1 << (2 * k)) - 1
    for canonical in (True, False):
        gt, ot = gpu_table(ctx, [s], k, canonical), ora_table([s], k, canonical)
        hi, lo, _ = ot.export()
        present = [(int(h) << 64) | int(l) for h, l in zip(hi[:1500], lo[:1500])]
        rc = []
        for key in present[:700]:
            r = 0
            for i in range(k):
                r = (r << 2) | (3 - ((key >> (2 * i)) & 3))
            rc.append(r)
        rand = [int.from_bytes(rng.bytes(16), "little") & mask for _ in range(1000)]
        probes = present + rc + rand + [mask, 0]
        p_hi = np.array([p >> 64 for p in probes], dtype=np.uint64)
        p_lo = np.array([p & (2**64 - 1) for p in probes], dtype=np.uint64)
        for canon_probe in (False, True):
            got = gt.lookup_wide(p_hi, p_lo, canon_probe)
            want = np.array([ot.lookup(x, canon_probe) for x in probes], dtype=np.uint64)
            assert np.array_equal(got, want), (k, canonical, canon_probe)
        with pytest.raises(kat_b200.KatGpuError):
            gt.lookup(p_lo)  # the one-word call refuses a wide table


def test_wide_add_pairs_and_filtered_export(ctx):
    k = 39
    s = mixed_stream(5)
    ot = ora_table([s], k)
    hi, lo, cn = ot.export()
    t = ctx.count_begin(k, True, 1 << 12)
    half = hi.size // 2
    t.add_pairs_wide(hi[:half], lo[:half], cn[:half]).add_pairs_wide(hi, lo, cn)  # the first half twice
    t.feed(np.frombuffer(b"ACGT" * 30, dtype=np.uint8)).end()
    want = O.Table(k, True).add(b"ACGT" * 30).add_pairs(lo[:half], cn[:half], hi[:half]).add_pairs(lo, cn, hi).finish()
    assert_same_multiset(t, want)
    whi, wlo, wcn = want.export()
    gc = np.zeros(wlo.size, dtype=np.uint32)
    for w in (whi, wlo):
        for b in range(32):
            c = (w >> np.uint64(2 * b)) & np.uint64(3)
            gc += ((c == 1) | (c == 2)).astype(np.uint32)
    for bounds in ((1, 10000, 0, 100), (2, 5, 12, 22), (3, 1 << 40, 0, k)):
        inb = (wcn >= bounds[0]) & (wcn <= bounds[1]) & (gc >= bounds[2]) & (gc <= bounds[3])
        for inside, sel in ((True, inb), (False, ~inb)):
            ghi, glo, gcn, n = t.export_wide(bounds, inside)
            assert n == int(sel.sum()) and np.array_equal(ghi, whi[sel]) and np.array_equal(glo, wlo[sel]) and np.array_equal(gcn, wcn[sel])


def test_wide_fastq_parsed_on_the_device(ctx):
    """katgpu_count_feed_fastq feeds the wide kernels too: raw 4-line FASTQ, parsed on the GPU, counted at k = 45."""
    from test_gpu_parity import _regular_fastq
    rng = np.random.default_rng(45)
    raw = _regular_fastq(rng, 20000, False, 4)
    stream, st = O.parse_file(raw, 45, 3)
    assert st == 0
    t = ctx.count_begin(45, True, 1 << 20)
    assert t.feed_fastq(raw, 3) is True
    t.end()
    assert_same_multiset(t, ora_table([stream], 45))


def test_wide_entry_points_take_one_word_tables(ctx):
    """The _wide calls accept k <= 32 tables (keys_hi ignored / zero), so a caller can use one code path for every k."""
    s = seqgen.genome_reads(20_000, 2000, 150, seed=8)
    gt, ot = gpu_table(ctx, [s], 27), ora_table([s], 27)
    assert_same_multiset(gt, ot)
    keys, counts = ot.export()
    assert np.array_equal(gt.lookup_wide(np.zeros(500, dtype=np.uint64), keys[:500]), counts[:500])
    with pytest.raises(kat_b200.KatGpuError):
        ctx.count_begin(65, True, 1 << 16)

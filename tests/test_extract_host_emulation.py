"""k = 1..32 (one-word keys): the extraction arithmetic of the benchmarked kernels — kat_b200/csrc/extract.cuh (pack4,
valid_starts16, Cut16's constant funnel shifts, Roll16, cut_kmer, mix_words) and table.cuh's hash bijection
(mix_key / unmix_key, revcomp_kmer) — compiled for the CPU as it is (tests/hostcheck/extract_emul.cc: the CUDA qualifiers
are empty outside nvcc, the integer intrinsics get host stand-ins) and run in bin_bases_kernel's own tiling, for EVERY
k from 1 to 32, canonical and not: the multiset of k-mers equals the oracle's, the rolling and the cut forms agree, the
hash is a bijection on 2k-bit integers that the walks invert.  On the GPU the same kernels are checked end to end
(tests/test_gpu_parity.py) for a subset of these k."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import seqgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "hostcheck", "_build", "extract_emul")


@pytest.fixture(scope="module", autouse=True)
def _build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck"), "_build/extract_emul"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def stream(seed, n_reads=260):
    """reads with errors and Ns, runs of one base, lower case, IUPAC codes, short records; ~45 KB = several 8 KiB tiles"""
    s = seqgen.genome_reads(3000, n_reads, 150, seed=seed).tobytes()
    return s + b"N" + b"T" * 300 + b"N" + b"a" * 200 + b"N" + b"acgtACGTRYacgt" * 20 + b"N" + b"GC" * 90 + b"NACGN" + s[:3000].lower() + b"\nACGTTGCA\r\nAC"


def emulate(tmp_path, k, canonical, data, n_starts=-1):
    p = tmp_path / "stream.bin"
    p.write_bytes(data)
    r = subprocess.run([EMUL, str(k), str(int(canonical)), str(p), str(n_starts)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    checks = dict(x.split("=") for x in lines[-1].split()[1:])
    keys = np.array([int(x) for x in lines[:-1]], dtype=np.uint64)
    return keys, {a: int(b) for a, b in checks.items()}


@pytest.mark.parametrize("canonical", [True, False])
@pytest.mark.parametrize("k", range(1, 33))
def test_emulated_extraction_equals_oracle(tmp_path, k, canonical):
    data = stream(200 + k)
    ot = O.Table.from_chunks([data], k, canonical)
    keys, checks = emulate(tmp_path, k, canonical, data)
    assert checks["roll_mismatch"] == 0 and checks["cut_mismatch"] == 0 and checks["mix_mismatch"] == 0, checks
    assert checks["keys"] == ot.total == len(keys)
    got_keys, got_counts = np.unique(keys, return_counts=True)
    want_keys, want_counts = ot.export()
    assert np.array_equal(got_keys, want_keys) and np.array_equal(got_counts.astype(np.uint64), want_counts)


@pytest.mark.parametrize("k", [5, 16, 17, 27, 32])
def test_emulated_extraction_ragged_ends(tmp_path, k):
    """Streams that end anywhere relative to the 8 KiB tile and its halo, and feeds whose k-mer STARTS end before their
    bytes do (katgpu_count_feed's chunks: n_starts < n_valid, the halo supplies the bases of the last k-mers)."""
    base = stream(7, n_reads=120)
    for n in (0, 1, k - 1, k, k + 1, 8191, 8192, 8193, 8192 + 31, 8192 + 32, 8192 + 33, 16384 - k, 16384 + k, len(base)):
        data = base[:n]
        keys, checks = emulate(tmp_path, k, True, data)
        ot = O.Table.from_chunks([data], k, True)
        assert checks["roll_mismatch"] == 0 and checks["cut_mismatch"] == 0 and len(keys) == ot.total, (k, n, checks, ot.total)
        assert np.array_equal(np.unique(keys), ot.export()[0]), (k, n)
    # starts limited: exactly the k-mers that start before n_starts
    data = base[:20000]
    for n_starts in (1, 15, 16, 17, 8192, 8200, 19990):
        keys, checks = emulate(tmp_path, k, False, data, n_starts)
        want = O.Table.from_chunks([data[:n_starts + k - 1]], k, False)
        assert len(keys) == want.total and np.array_equal(np.unique(keys), want.export()[0]), (k, n_starts)

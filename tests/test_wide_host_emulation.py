"""k = 33..64 (two-word keys): the device code of kat_b200/csrc/wide.cuh — rolling extraction in the kernel's own
tiling, canonicalisation, hashing, the insert / lookup protocol over the wide table, growth, reverse complement, GC
count — is __host__ __device__; here it runs on the CPU (tests/hostcheck/wide_emul.cc, atomics replaced by
single-threaded stand-ins) and is compared with the oracle.  The same checks run against the real kernels in
tests/test_gpu_wide.py."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
import seqgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "hostcheck", "_build", "wide_emul")


@pytest.fixture(scope="module", autouse=True)
def _build():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck"), "_build/wide_emul"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def stream(seed):
    """reads with errors and Ns, lower case, a long poly-T and poly-A run (the all-ones key of k = 64), short records"""
    s = seqgen.genome_reads(3000, 400, 150, seed=seed).tobytes()
    return s + b"N" + b"T" * 300 + b"N" + b"a" * 200 + b"N" + b"acgtACGTRYacgt" * 20 + b"N" + b"GC" * 90 + b"N" + s[:3000].lower()


def run_emul(tmp_path, k, canonical, slots_log2, data, chunk, probes=None, canon_probe=0):
    p = tmp_path / "stream.bin"
    p.write_bytes(data)
    args = [EMUL, str(k), str(int(canonical)), str(slots_log2), str(p), str(chunk)]
    if probes is not None:
        q = tmp_path / "probes.txt"
        q.write_text("".join(f"{int(x) >> 64} {int(x) & (2**64 - 1)}\n" for x in probes))
        args += [str(q), str(canon_probe)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.splitlines()
    stats = [int(x) for x in lines[-1].split()[1:]]
    return lines[:-1], stats


@pytest.mark.parametrize("k", [33, 40, 48, 63, 64])
@pytest.mark.parametrize("canonical", [True, False])
def test_emulated_count_equals_oracle(tmp_path, k, canonical):
    data = stream(100 + k)
    ot = O.Table.from_chunks([data], k, canonical)
    hi, lo, cnt = ot.export()
    want = sorted(zip(hi.tolist(), lo.tolist(), cnt.tolist()))
    # one chunk; a table that has to grow several times; chunks of odd sizes (halo / ragged tiles / tiny last chunk)
    for slots_log2, chunk in ((20, 1 << 30), (10, 1 << 30), (12, 9216 * 2), (12, 5003), (14, len(data) - 7)):
        lines, (claimed, total, failed, grows) = run_emul(tmp_path, k, canonical, slots_log2, data, chunk)
        if chunk >= len(data):  # chunking by start position never changes the multiset; check the others too
            pass
        got = sorted(tuple(int(x) for x in l.split()) for l in lines)
        assert got == want, (k, canonical, slots_log2, chunk)
        sentinel = int(ot.lookup((1 << 128) - 1)) if k == 64 else 0
        assert failed == 0 and total == ot.total and claimed == ot.distinct - (1 if sentinel else 0)
        if slots_log2 == 10:
            assert grows >= 3
    if k == 64 and not canonical:
        assert ot.lookup((1 << 128) - 1) == 300 - 64 + 1  # T x 64: the key that equals the empty marker


@pytest.mark.parametrize("k", [33, 47, 64])
def test_emulated_lookup_revcomp_gc(tmp_path, k):
    data = stream(7)
    rng = np.random.default_rng(k)
    for canonical in (True, False):
        ot = O.Table.from_chunks([data], k, canonical)
        hi, lo, _ = ot.export()
        present = [(int(h) << 64) | int(l) for h, l in zip(hi[:300], lo[:300])]
        rc = []
        for key in present[:150]:  # reverse complements of stored keys: found only when the probe is canonicalised
            r = 0
            for i in range(k):
                r = (r << 2) | (3 - ((key >> (2 * i)) & 3))
            rc.append(r)
        rand = [int.from_bytes(rng.bytes(16), "little") & ((1 << (2 * k)) - 1) for _ in range(200)]
        probes = present + rc + rand + [(1 << (2 * k)) - 1, 0]
        for canon_probe in (0, 1):
            lines, _ = run_emul(tmp_path, k, canonical, 16, data, 1 << 30, probes, canon_probe)
            got = [int(l.split()[1]) for l in lines]
            want = [ot.lookup(x, bool(canon_probe)) for x in probes]
            assert got == want, (k, canonical, canon_probe)
            gcs = [int(l.split()[2]) for l in lines]
            assert gcs == [sum(((x >> (2 * i)) & 3) in (1, 2) for i in range(k)) for x in probes]

"""CPU-only checks of the product's boundary: the C-ABI library loads and exports
every symbol include/katgpu.h declares, refuses to run without a GPU (no CPU
fallback), and the C++ front-end's FASTA/FASTQ reader emits the reference's base
stream (checked against the oracle on the golden inputs and on random files)."""
import os
import re
import subprocess

import numpy as np
import pytest

import golden_util as G
import oracle as O
import kat_b200
from kat_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = os.path.join(ROOT, "kat_b200", "bin", "kat")


def test_every_declared_symbol_is_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "katgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(katgpu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = B.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), name + " declared in katgpu.h but not exported by libkatgpu.so"
    assert declared == set(B.SIGNATURES), declared ^ set(B.SIGNATURES)
    assert lib.katgpu_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert B.lib().katgpu_device_count() == 0
    with pytest.raises(kat_b200.KatGpuError) as e:
        kat_b200.Context(0)
    assert e.value.status == -5 and "no CPU path" in str(e.value)
    r = subprocess.run([KAT, "hist", "-m5", os.path.join(G.GOLDEN, "inputs", "edge1.fa")], capture_output=True, text=True)
    assert r.returncode == 4 and "no CPU path" in r.stderr


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, link or exec anything under oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "kat_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "libkatoracle" not in txt and "import oracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), f
    out = subprocess.run(["ldd", B.lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out


def _kat_parse(path, k, trim=0, want_route=None, carry=None):
    """carry: one-element list with the seq_len state inherited from the previous file of a group (updated in place)."""
    r = subprocess.run([KAT, "_parse", "-m", str(k), "--5ptrim", str(trim), "--carry", str(carry[0] if carry else 0), path], capture_output=True,
                       env=dict(os.environ, KAT_TRACE="1"))
    if want_route is not None:
        assert b"parser: " + want_route in r.stderr, r.stderr
    if carry is not None and b"carry_out=" in r.stderr:
        carry[0] = int(r.stderr.split(b"carry_out=")[1].split()[0])
    return r.stdout, {0: 0, 2: -1, 3: -2}[r.returncode]


@pytest.mark.parametrize("name", sorted(os.listdir(os.path.join(G.GOLDEN, "inputs"))))
@pytest.mark.parametrize("k,trim", [(5, 0), (21, 0), (27, 3), (32, 7)])
def test_host_parser_matches_oracle_on_golden_inputs(name, k, trim):
    import gzip
    path = os.path.join(G.GOLDEN, "inputs", name)
    data = open(path, "rb").read()
    if data[:2] == b"\x1f\x8b":
        data = gzip.decompress(data)
    want, st = O.parse_file(data, k, trim)
    got, gst = _kat_parse(path, k, trim)
    if st == -2:
        assert gst == -2  # the front-end reports it; what was emitted before matches
    assert (got, gst) == (want, st)


def _random_file(rng, fastq):
    out = bytearray()
    for i in range(int(rng.integers(1, 40))):
        n = int(rng.integers(0, 9000)) if rng.random() < 0.15 else int(rng.integers(0, 300))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTacgtNRY-", dtype=np.uint8), n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .01, .01]))
        width = int(rng.integers(1, 200)) if rng.random() < 0.5 else max(n, 1)
        lines = [seq[j:j + width] for j in range(0, n, width)] or [b""]
        if fastq:
            out += b"@r%d some text\n" % i + b"\n".join(lines) + b"\n+\n"
            q = bytes(rng.choice(np.frombuffer(b"I@+>#5", dtype=np.uint8), n))
            out += b"\n".join(q[j:j + width] for j in range(0, n, width)) + b"\n"
        else:
            out += b">c%d\n" % i + b"\n".join(lines) + b"\n"
            if rng.random() < 0.2:
                out += b"\n"
    if rng.random() < 0.3 and out.endswith(b"\n"):
        out = out[:-1]
    return bytes(out)


@pytest.mark.parametrize("seed", range(24))
def test_host_parser_matches_oracle_on_random_files(seed, tmp_path):
    rng = np.random.default_rng(seed)
    data = _random_file(rng, fastq=seed % 2 == 0)
    p = tmp_path / ("r.fq" if seed % 2 == 0 else "r.fa")
    p.write_bytes(data)
    for k, trim in [(5, 0), (27, 0), (21, int(rng.integers(1, 12)))]:
        assert _kat_parse(str(p), k, trim) == O.parse_file(data, k, trim), (seed, k, trim)


@pytest.mark.parametrize("seed", range(4))
def test_host_parser_matches_oracle_on_token_soup(seed, tmp_path):
    """Malformed input: 75 random mixtures of record markers, line ends, bases and quality-like bytes per seed; the host
    parser emits the oracle's stream and status for each (the oracle itself is held to the reference binary on such
    files by tests/test_oracle_pinning.py::test_oracle_vs_reference_binary_on_token_soup)."""
    import seqgen
    rng = np.random.default_rng(1000 + seed)
    p = tmp_path / "soup.fq"
    for it in range(75):
        data = seqgen.token_soup(rng, always_marker=False)
        p.write_bytes(data)
        for k, trim in ((5, 0), (3, 2)):
            assert _kat_parse(str(p), k, trim) == O.parse_file(data, k, trim), (seed, it, k, trim, data)


@pytest.mark.parametrize("seed", range(3))
def test_host_parser_carries_seq_len_across_the_files_of_a_group(seed, tmp_path):
    """The reference parser keeps `seq_len` when a stream moves to its next file (JFI/mer_overlap_sequence_parser.hpp:158-187):
    a file that ends inside a FASTQ record makes the next FASTQ file's first record "too long" and the rest of the group is
    dropped.  Groups of 2-4 files (soups, regular files, files with an empty or truncated last record): the host parser,
    fed file by file with the carried state, emits the oracle's streams, statuses and final state."""
    import seqgen
    rng = np.random.default_rng(3000 + seed)
    tails = [b"@t\n\n+\n\n", b"@t\nACGTACGT\n", b"@t\nACGTAC", b"@t\nACGTACGT\n+\nIIIIIIII\n", b""]
    for it in range(25):
        files = []
        for j in range(int(rng.integers(2, 5))):
            kind = rng.random()
            body = seqgen.token_soup(rng) if kind < 0.4 else _regular_file(rng, kind < 0.8, int(rng.integers(1, 6)))
            files.append(body + (tails[int(rng.integers(0, len(tails)))] if body.startswith(b"@") else b""))
        host, ora = [0], [0]
        for j, data in enumerate(files):
            p = tmp_path / ("g%d_%d.fq" % (it, j))
            p.write_bytes(data)
            got = _kat_parse(str(p), 5, 0, carry=host)
            want = O.parse_file(data, 5, 0, ora)
            assert got == want and (want[1] != 0 or host == ora), (seed, it, j, data[-40:], host, ora)
            if want[1] != 0:
                break


# ---- the parser's fast paths (regular 4-line FASTQ / FASTA layouts, parsed piece-wise in parallel)
def _regular_file(rng, fastq, n_records, crlf=False, tricky_quals=True, no_final_newline=False, wrap=0):
    nl = b"\r\n" if crlf else b"\n"
    out = bytearray()
    for i in range(n_records):
        n = int(rng.integers(12, 400)) if rng.random() < 0.97 else int(rng.integers(4000, 9000))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTacgtNRY-", dtype=np.uint8), n, p=[.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .01, .01]))
        if fastq:
            q = bytes(rng.choice(np.frombuffer(b"I@+>#5" if tricky_quals else b"IJK#5", dtype=np.uint8), n))
            plus = b"+" if rng.random() < 0.7 else b"+r%d again" % i
            out += b"@r%d/1 x@y+z" % i + nl + seq + nl + plus + nl + q + nl
        else:
            w = wrap or n
            out += b">ctg%d len=%d" % (i, n) + nl + nl.join(seq[j:j + w] for j in range(0, n, w)) + nl
    if no_final_newline:
        out = out[:-len(nl)]
    return bytes(out)


@pytest.mark.parametrize("fastq,kw", [(True, {}), (True, {"crlf": True}), (True, {"no_final_newline": True}),
                                      (False, {}), (False, {"wrap": 70}), (False, {"wrap": 61, "no_final_newline": True}),
                                      (False, {"wrap": 70, "crlf": True})],
                         ids=["fq", "fq_crlf", "fq_no_final_nl", "fa", "fa_wrapped", "fa_wrapped_no_final_nl", "fa_crlf"])
@pytest.mark.parametrize("piece,threads", [(700, 3), (5000, 1), (1 << 24, 4)])
def test_host_parser_fast_paths_equal_the_emulation(fastq, kw, piece, threads, tmp_path, monkeypatch):
    """Regular layouts go through the memchr fast path, cut in pieces and parsed by several
    threads; the result must be the byte stream of the character-level emulation of the
    reference parser (forced with KAT_PARSE_EXACT_ONLY) and of the oracle."""
    rng = np.random.default_rng(piece + threads + len(kw) + fastq)
    data = _regular_file(rng, fastq, 600, **kw)
    p = tmp_path / ("r.fq" if fastq else "r.fa")
    p.write_bytes(data)
    for k, trim in [(27, 0), (21, 5), (5, 11)]:
        monkeypatch.setenv("KAT_PARSE_PIECE_BYTES", str(piece))
        monkeypatch.setenv("KAT_PARSE_THREADS", str(threads))
        monkeypatch.delenv("KAT_PARSE_EXACT_ONLY", raising=False)
        # FASTA with --5ptrim keeps the reference's per-buffer quirk: emulation only; a FASTQ whose
        # last line lacks its newline is "Invalid fastq" territory: emulation only
        expect_fast = not (not fastq and trim) and not (fastq and kw.get("no_final_newline"))
        fast = _kat_parse(str(p), k, trim, b"fast" if expect_fast else b"exact")
        monkeypatch.setenv("KAT_PARSE_EXACT_ONLY", "1")
        exact = _kat_parse(str(p), k, trim, b"exact")
        assert fast == exact == O.parse_file(data, k, trim), (k, trim)


@pytest.mark.parametrize("mutate", ["blank_line", "multi_line_seq", "plus_in_seq", "short_qual", "seq_shorter_than_trim",
                                    "gt_in_fasta_seq", "at_seq", "trailing_garbage"])
def test_host_parser_irregular_inputs_fall_back(mutate, tmp_path, monkeypatch):
    """Anything outside the regular layouts must come out exactly as the emulation gives it."""
    rng = np.random.default_rng(len(mutate))
    fastq = mutate != "gt_in_fasta_seq"
    recs = _regular_file(rng, fastq, 40, tricky_quals=False).split(b"\n")
    i = 4 * 17 + 1 if fastq else 2 * 17 + 1  # the sequence line of record 17
    if mutate == "blank_line":
        recs.insert(i - 1, b"")
    elif mutate == "multi_line_seq":
        s = recs[i]
        recs[i:i + 1] = [s[:5], s[5:]]
        q = recs[i + 3]
        recs[i + 3:i + 4] = [q[:7], q[7:]]
    elif mutate == "plus_in_seq":
        recs[i] = recs[i][:3] + b"+" + recs[i][4:]
    elif mutate == "short_qual":
        recs[i + 2] = recs[i + 2][:-2]
    elif mutate == "seq_shorter_than_trim":
        recs[i] = recs[i][:3]
        recs[i + 2] = recs[i + 2][:3]
    elif mutate == "gt_in_fasta_seq":
        recs[i] = recs[i][:6] + b">" + recs[i][7:]
    elif mutate == "at_seq":
        recs[i] = b"@" + recs[i][1:]
    elif mutate == "trailing_garbage":
        recs.append(b"garbage without newline")
    data = b"\n".join(recs)
    p = tmp_path / ("m.fq" if fastq else "m.fa")
    p.write_bytes(data)
    monkeypatch.setenv("KAT_PARSE_PIECE_BYTES", "900")
    for k, trim in [(21, 0), (21, 6)]:
        monkeypatch.delenv("KAT_PARSE_EXACT_ONLY", raising=False)
        fast = _kat_parse(str(p), k, trim, b"exact")  # the fast path must have refused the file
        monkeypatch.setenv("KAT_PARSE_EXACT_ONLY", "1")
        assert fast == _kat_parse(str(p), k, trim) == O.parse_file(data, k, trim), (mutate, k, trim)


@pytest.mark.parametrize("fastq,kw", [(True, {}), (True, {"crlf": True}), (False, {}), (False, {"wrap": 70}), (False, {"wrap": 61, "no_final_newline": True})],
                         ids=["fq", "fq_crlf", "fa", "fa_wrapped", "fa_wrapped_no_final_nl"])
@pytest.mark.parametrize("block", [300, 4096, 70000])
def test_host_streaming_route_equals_whole_file(fastq, kw, block, tmp_path, monkeypatch):
    """Large regular files are read, parsed and fed block by block (stream_regular_file); the blocks,
    joined by the 'N' that separates records, are the whole-file base stream."""
    rng = np.random.default_rng(block + fastq)
    data = _regular_file(rng, fastq, 500, **kw)
    p = tmp_path / ("s.fq" if fastq else "s.fa")
    p.write_bytes(data)
    monkeypatch.setenv("KAT_STREAM_BYTES", "1")
    monkeypatch.setenv("KAT_STREAM_BLOCK", str(block))
    for k, trim in [(27, 0), (21, 4)]:
        r = subprocess.run([KAT, "_parse", "-m", str(k), "--5ptrim", str(trim), str(p)], capture_output=True, env=dict(os.environ, KAT_TRACE="1"))
        if fastq or trim == 0:
            assert b"stream route: rc 1" in r.stderr, r.stderr
        assert (r.stdout, r.returncode) == (O.parse_file(data, k, trim)[0], 0), (k, trim)


@pytest.mark.parametrize("mutate_at", [3, 250, 499])
def test_host_streaming_route_gives_up_on_irregular_files(mutate_at, tmp_path, monkeypatch):
    """An irregular record anywhere makes the streaming route report it (0 = nothing emitted yet,
    -1 = start over); the hook then takes the whole-file route and the stream is the emulation's."""
    rng = np.random.default_rng(mutate_at)
    recs = _regular_file(rng, True, 500, tricky_quals=False).split(b"\n")
    s = recs[4 * mutate_at + 1]
    recs[4 * mutate_at + 1:4 * mutate_at + 2] = [s[:6], s[6:]]  # a two-line sequence
    data = b"\n".join(recs)
    p = tmp_path / "m.fq"
    p.write_bytes(data)
    monkeypatch.setenv("KAT_STREAM_BYTES", "1")
    monkeypatch.setenv("KAT_STREAM_BLOCK", "5000")
    r = subprocess.run([KAT, "_parse", "-m", "21", str(p)], capture_output=True, env=dict(os.environ, KAT_TRACE="1"))
    assert b"stream route: rc 1" not in r.stderr
    want, st = O.parse_file(data, 21, 0)
    assert r.stdout == want


def test_raw_fastq_blocks_are_whole_records(tmp_path):
    """stream_fastq_raw (the host side of the device FASTQ parser): a file is handed out in blocks, and every
    block in per-GPU shares, that start at a record and end with a newline, are contiguous and cover the file once —
    also when quality lines start with '@' or '+'.  A file that is not FASTQ or lacks the final newline is declined."""
    rng = np.random.default_rng(3)
    recs = []
    for i in range(3000):
        L = int(rng.integers(20, 200))
        seq = bytes(rng.choice(list(b"ACGTN"), L).tolist())
        q = bytes(rng.choice(list(b"@+I#"), L).tolist())
        recs.append(b"@r%d some text\n" % i + seq + b"\n+\n" + q + b"\n")
    data = b"".join(recs)
    p = tmp_path / "r.fq"
    p.write_bytes(data)
    kat = os.path.join(ROOT, "kat_b200", "bin", "kat")
    for block, parts in ((4096, 1), (50_000, 3), (10_000_000, 8), (7777, 2)):
        out = subprocess.run([kat, "_rawsplit", str(p), str(block), str(parts)], capture_output=True, text=True, check=True).stdout
        assert "!" not in out and "?" not in out, out[-400:]
        assert out.strip().splitlines()[-1] == "rc 1 total %d" % len(data)
        if block < len(data):
            assert out.count("block ") > 1
    (tmp_path / "nonl.fq").write_bytes(data[:-1])
    (tmp_path / "x.fa").write_bytes(b">a\nACGT\n")
    for name in ("nonl.fq", "x.fa"):
        out = subprocess.run([kat, "_rawsplit", str(tmp_path / name), "4096", "2"], capture_output=True, text=True, check=True).stdout
        assert out.strip().splitlines()[-1].startswith("rc 0")

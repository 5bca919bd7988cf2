"""The product front-end (`kat_b200/bin/kat hist|gcp|comp`, C++ over libkatgpu) must
write the same FILES as the reference binary: every golden case is re-run with the
product, in a copy of tests/golden, and every output file is compared byte for byte
(headers, matrices, .stats including the printed doubles, .N.hist)."""
import os
import shutil
import subprocess

import pytest

import golden_util as G

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = os.path.join(ROOT, "kat_b200", "bin", "kat")


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_run")
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), d / "inputs")
    return d


@pytest.mark.parametrize("path", ["direct", "partition", "tiled"])
@pytest.mark.parametrize("case", G.cases(), ids=lambda c: c["name"])
def test_cli_output_files_byte_identical(case, path, workdir):
    out = os.path.join("expected", case["name"])  # same relative prefix as the reference run -> same headers
    shutil.rmtree(workdir / out, ignore_errors=True)
    os.makedirs(workdir / out)
    args = [a.replace("OUT", out) for a in case["args"]]
    env = dict(os.environ, KATGPU_COUNT_PATH=path)
    r = subprocess.run([KAT] + args, cwd=workdir, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    produced = sorted(os.listdir(workdir / out))
    assert produced == case["files"]
    for f in case["files"]:
        got = open(workdir / out / f, "rb").read()
        want = G.expected_bytes(case["name"], f)
        assert got == want, f"{case['name']}/{f} differs from the reference's file"


@pytest.mark.parametrize("path", ["auto", "tiled"])
def test_cli_c1_ecoli_k27(path, tmp_path):
    """BASELINE configs[0] through the PRODUCT on a GPU: `kat hist -m27` on the reference's own test
    genome (tests/golden/c1/EcoliK12.fasta.gz) writes the reference's file byte for byte (bins sum to
    4,550,880 distinct / 4,641,626 instances, tests/check_jellyfish.cc:118-135), and
    `kat comp -m13 reads genome` writes its matrix and .stats byte for byte."""
    import gzip
    shutil.copytree(os.path.join(G.GOLDEN, "c1"), tmp_path / "c1", ignore=shutil.ignore_patterns("out"))
    os.makedirs(tmp_path / "c1" / "out")
    env = dict(os.environ)
    if path != "auto":
        env["KATGPU_COUNT_PATH"] = path
    for args in (["hist", "-m27", "-t1", "-o", "c1/out/ecoli.hist", "c1/EcoliK12.fasta.gz"],
                 ["comp", "-m13", "-t1", "-o", "c1/out/r1_vs_k12", "c1/ecoli_r1.1K.fastq.gz", "c1/EcoliK12.fasta.gz"]):
        r = subprocess.run([KAT] + args, cwd=tmp_path, capture_output=True, text=True, env=env, timeout=300)
        assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    for f in ("ecoli.hist", "r1_vs_k12-main.mx", "r1_vs_k12.stats"):
        want = gzip.open(os.path.join(G.GOLDEN, "c1", "out", f + ".gz"), "rb").read()
        assert open(tmp_path / "c1" / "out" / f, "rb").read() == want, f + " differs from the reference's file"
    lab, val = G.parse_hist_text(open(tmp_path / "c1" / "out" / "ecoli.hist", "rb").read())
    assert int(val.sum()) == 4550880 and int((lab * val).sum()) == 4641626


def test_cli_hash_full_exit_code(workdir):
    """--disable_hash_grow + a table that is too small: the reference throws
    std::runtime_error("Hash full") -> "Error: Hash full", exit code 5 (src/kat.cc:292-294)."""
    r = subprocess.run([KAT, "comp", "-m21", "-g", "-H", "1000", "-I", "1000", "-o", "hf/c", "inputs/synth_a.fq", "inputs/synth_c.fa"],
                       cwd=workdir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 5 and "Hash full" in r.stderr


def test_cli_missing_input(workdir):
    r = subprocess.run([KAT, "hist", "-m21", "inputs/nope.fa"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 4 and "Could not find input file" in r.stderr


# ---- jellyfish hashes either side of the path: LOAD mode and -d/--dump_hash -------------------
import gzip
import json

JF_MANIFEST = json.load(open(os.path.join(G.GOLDEN, "jf", "cases.json")))
REF_JF = os.path.join(ROOT, "oracle", "_ref", "kat_jellyfish")


@pytest.fixture(scope="module")
def jf_workdir(tmp_path_factory):
    d = tmp_path_factory.mktemp("jf_run")
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), d / "inputs")
    os.makedirs(d / "jf")
    for h in JF_MANIFEST["hashes"]:
        shutil.copy(os.path.join(G.GOLDEN, "jf", h["name"] + ".jf"), d / "jf")
    return d


def _jf_expected(case, f):
    p = os.path.join(G.GOLDEN, "jf", "expected", case, f)
    return open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()


@pytest.mark.parametrize("case", JF_MANIFEST["cases"], ids=lambda c: c["name"])
def test_cli_load_mode_files_byte_identical(case, jf_workdir):
    """Hashes written by the reference (its own -d, and jellyfish count) as INPUT of the product's
    hist / gcp / comp (InputHandler::loadHash): same output files as the reference, byte for byte
    — including k taken from the hash header and the canonical flag of a loaded input."""
    out = os.path.join("jf", "expected", case["name"])
    shutil.rmtree(jf_workdir / out, ignore_errors=True)
    os.makedirs(jf_workdir / out)
    r = subprocess.run([KAT] + [a.replace("OUT", out) for a in case["args"]], cwd=jf_workdir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert "Loading hashes into memory..." in r.stdout
    for f in case["files"]:
        assert open(jf_workdir / out / f, "rb").read() == _jf_expected(case["name"], f), f"{case['name']}/{f}"


@pytest.mark.parametrize("path", ["direct", "partition", "tiled"])
def test_cli_dump_hash_is_a_jellyfish_file(path, jf_workdir):
    """`kat hist -d` writes <prefix>-hash.jf<k>: the same (k-mer, count) records as the reference's
    own dump of the same input, readable by the unmodified jellyfish, and loadable again."""
    env = dict(os.environ, KATGPU_COUNT_PATH=path)
    r = subprocess.run([KAT, "hist", "-m13", "-d", "-o", "dump/h", "inputs/long.fa"], cwd=jf_workdir, capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "Dumping hash to dump/h-hash.jf13 ..." in r.stdout, r.stdout + r.stderr
    want = sorted(gzip.open(os.path.join(G.GOLDEN, "jf", "long_k13_kat.dump.txt.gz"), "rt").read().splitlines())
    d = subprocess.run([KAT, "_jfdump", "dump/h-hash.jf13"], cwd=jf_workdir, capture_output=True, text=True)
    lines = d.stdout.splitlines()
    assert "canonical=1 counter_len=4" in lines[0] and lines[-1].endswith("out_of_order=0")
    assert sorted(l for l in lines if not l.startswith("#")) == want
    if os.access(REF_JF, os.X_OK):
        j = subprocess.run([REF_JF, "dump", "-c", "dump/h-hash.jf13"], cwd=jf_workdir, capture_output=True, text=True)
        assert j.returncode == 0 and sorted(j.stdout.splitlines()) == want
    # and our own LOAD of our own dump reproduces the histogram of the counting run
    r2 = subprocess.run([KAT, "hist", "-o", "dump/h2", "dump/h-hash.jf13"], cwd=jf_workdir, capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stderr
    body = lambda p: [l for l in open(jf_workdir / p).read().splitlines() if not l.startswith("#")]
    assert body("dump/h") == body("dump/h2")


def test_cli_comp_dump_hashes_and_symlink(jf_workdir):
    """comp -d: counted inputs are dumped as <prefix>-hash<i>.jf<k>, loaded ones symlinked
    (InputHandler::dump, lib/src/input_handler.cc:221-243)."""
    r = subprocess.run([KAT, "comp", "-m21", "-d", "-i", "40", "-j", "50", "-o", "cd/c", "jf/synthc_k21_jf.jf", "inputs/synth_a.fq"],
                       cwd=jf_workdir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert os.path.islink(jf_workdir / "cd" / "c-hash1.jf21")
    d = subprocess.run([KAT, "_jfdump", "cd/c-hash2.jf21"], cwd=jf_workdir, capture_output=True, text=True)
    assert d.returncode == 0 and "k=21 canonical=1" in d.stdout.splitlines()[0]
    assert open(jf_workdir / "cd" / "c-main.mx", "rb").read() == _jf_expected("load_comp_hash_vs_reads_k21", "c-main.mx")


@pytest.mark.skipif(not os.path.isdir("/dev/shm"), reason="no /dev/shm")
def test_cli_dev_paths_are_pipes(workdir):
    """JellyfishHelper::isPipe: every path under /dev or /proc — a regular file in /dev/shm too —
    is printed as <pipe> in the output headers (InputHandler::pathString, lib/src/input_handler.cc:160-170)."""
    import tempfile
    d = tempfile.mkdtemp(prefix="katcli_", dir="/dev/shm")
    try:
        shutil.copy(workdir / "inputs" / "synth_c.fa", os.path.join(d, "c.fa"))
        r = subprocess.run([KAT, "hist", "-m21", "-o", os.path.join(d, "h"), os.path.join(d, "c.fa")], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        head = open(os.path.join(d, "h")).read().splitlines()[:5]
        assert head[0] == "# Title:21-mer spectra for: c.fa" and head[4] == "# Input 1:<pipe>"
    finally:
        shutil.rmtree(d, ignore_errors=True)


# ---- kat sect: per-position coverage through batched GPU lookups ------
SECT_MANIFEST = json.load(open(os.path.join(G.GOLDEN, "sect", "cases.json")))


@pytest.mark.parametrize("case", SECT_MANIFEST, ids=lambda c: c["name"])
def test_cli_sect_and_cold_output_files_byte_identical(case, tmp_path):
    """`kat sect` (src/sect.cc) and `kat cold` (src/cold.cc): counts.cvg, stats.tsv, counts.gc and the region FASTA files equal the
    files written by the reference binary (tests/golden/sect, make_sect_golden.py)."""
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), tmp_path / "inputs")
    out = os.path.join("sect", case["name"])
    os.makedirs(tmp_path / out)
    args = [a.replace("OUT", out) for a in case["args"]]
    r = subprocess.run([KAT] + args, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert sorted(os.listdir(tmp_path / out)) == case["files"]
    for f in case["files"]:
        p = os.path.join(G.GOLDEN, "sect", case["name"], f)
        want = open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()
        assert open(tmp_path / out / f, "rb").read() == want, f"{case['name']}/{f} differs from the reference's file"


# ---- kat filter seq / kat filter kmer ----------------------------------
FILTER_MANIFEST = json.load(open(os.path.join(G.GOLDEN, "filter", "cases.json")))


def _filter_expected(case, f):
    p = os.path.join(G.GOLDEN, "filter", case, f)
    return open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()


@pytest.mark.parametrize("case", FILTER_MANIFEST, ids=lambda c: c["name"])
def test_cli_filter_matches_reference(case, tmp_path):
    """`kat filter seq` (src/filter_sequence.cc): the kept / discarded record files and the .stats table are the
    reference's files byte for byte.  `kat filter kmer` (src/filter_kmer.cc): the hashes written hold exactly the
    reference's (k-mer, count) records, in jellyfish's sorted order, and the printed counters are the same
    (tests/golden/filter, make_filter_golden.py)."""
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), tmp_path / "inputs")
    os.makedirs(tmp_path / "jf")
    shutil.copy(os.path.join(G.GOLDEN, "jf", "synthc_k21_jf.jf"), tmp_path / "jf")
    out = os.path.join("filter", case["name"])
    os.makedirs(tmp_path / out)
    args = [a.replace("OUT", out) for a in case["args"]]
    r = subprocess.run([KAT] + args, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert sorted(os.listdir(tmp_path / out)) == case["files"]
    assert [l for l in r.stdout.splitlines() if l.startswith(("K-mers ", "Found "))] == case["summary"]
    for f in case["files"]:
        if ".jf" in f:
            d = subprocess.run([KAT, "_jfdump", os.path.join(out, f)], cwd=tmp_path, capture_output=True, text=True)
            lines = d.stdout.splitlines()
            assert d.returncode == 0 and lines[-1].endswith("out_of_order=0"), d.stdout[-300:] + d.stderr
            want = _filter_expected(case["name"], f + ".dump.txt").decode().splitlines()
            assert sorted(l for l in lines if not l.startswith("#")) == want, f"{case['name']}/{f}: records differ"
        else:
            assert open(tmp_path / out / f, "rb").read() == _filter_expected(case["name"], f), f"{case['name']}/{f} differs from the reference's file"


def test_cli_filter_errors(tmp_path):
    """Argument errors of the filter tools: exit code 4 with the reference's messages."""
    fa = os.path.join(G.GOLDEN, "inputs", "edge1.fa")
    for args, msg in ((["filter", "kmer", "-c", "5", "-d", "2", fa], "High kmer count value must be >= to low kmer count value"),
                      (["filter", "kmer", "-g", "9", "-h", "2", fa], "High GC count value must be >= to low GC count value"),
                      (["filter", "seq", "-m5", fa], "You must specify at least one sequence file to filter"),
                      (["filter", "seq", "-m5", "--seq", "nope.fa", fa], "Could not find input file at: nope.fa"),
                      (["filter", "nonsense", fa], "Could not recognise mode string: nonsense")):
        r = subprocess.run([KAT] + args, cwd=tmp_path, capture_output=True, text=True, timeout=120)
        assert r.returncode == 4 and msg in r.stderr, (args, r.returncode, r.stderr)


# ---- mer lengths above 32: two-word keys ------------------------------
WIDE_MANIFEST = json.load(open(os.path.join(G.GOLDEN, "wide", "cases.json")))


@pytest.mark.parametrize("case", WIDE_MANIFEST, ids=lambda c: c["name"])
def test_cli_wide_kmers_output_files_byte_identical(case, tmp_path):
    """hist / gcp / comp (2 and 3 inputs, mixed canonical) / sect / cold / filter seq at k = 33..64: every output file equals
    the file the reference binary writes (tests/golden/wide, make_wide_golden.py)."""
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), tmp_path / "inputs")
    out = os.path.join("wide", case["name"])
    os.makedirs(tmp_path / out)
    args = [a.replace("OUT", out) for a in case["args"]]
    r = subprocess.run([KAT] + args, cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert sorted(os.listdir(tmp_path / out)) == case["files"]
    for f in case["files"]:
        p = os.path.join(G.GOLDEN, "wide", case["name"], f)
        want = open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()
        assert open(tmp_path / out / f, "rb").read() == want, f"{case['name']}/{f} differs from the reference's file"


def test_cli_wide_kmers_limits(tmp_path):
    """Mer lengths above 64 are refused with a message (two 64-bit words per k-mer at most); a mer length of 0 fails the way the
    reference does (a hash matrix without columns: "Invalid matrix size 27x0", exit code 5)."""
    fa = os.path.join(G.GOLDEN, "inputs", "long.fa")
    for args, code, msg in ((["hist", "-m65", fa], 4, "mer lengths 1..64"), (["comp", "-m0", fa, fa], 5, "Invalid matrix size 27x0"),
                            (["sect", "-m100", fa, fa], 4, "mer lengths 1..64")):
        r = subprocess.run([KAT] + args, cwd=tmp_path, capture_output=True, text=True, timeout=120)
        assert r.returncode == code and msg in r.stderr, (args, r.returncode, r.stderr)


WIDE_HASH_MANIFEST = json.load(open(os.path.join(G.GOLDEN, "wide", "hash_cases.json")))


@pytest.mark.parametrize("case", WIDE_HASH_MANIFEST, ids=lambda c: c["name"])
def test_cli_wide_kmers_hash_files(case, tmp_path):
    """Jellyfish hashes with two-word keys (k > 32) either side of the path: a hash written by the reference is loaded
    (hist / gcp / comp outputs byte-identical to the reference's own LOAD run), and the hashes the product writes
    (-d dumps, filter kmer) hold exactly the records of the reference's, in jellyfish's sorted order."""
    shutil.copytree(os.path.join(G.GOLDEN, "inputs"), tmp_path / "inputs")
    os.makedirs(tmp_path / "wide")
    shutil.copy(os.path.join(G.GOLDEN, "wide", "ref_small_k41.jf"), tmp_path / "wide")
    out = os.path.join("wide", case["name"])
    os.makedirs(tmp_path / out)
    r = subprocess.run([KAT] + [a.replace("OUT", out) for a in case["args"]], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
    assert sorted(os.listdir(tmp_path / out)) == case["files"]
    for f in case["files"]:
        p = os.path.join(G.GOLDEN, "wide", case["name"], f)
        if ".jf" in f:
            d = subprocess.run([KAT, "_jfdump", os.path.join(out, f)], cwd=tmp_path, capture_output=True, text=True)
            lines = d.stdout.splitlines()
            assert d.returncode == 0 and lines[-1].endswith("out_of_order=0"), d.stdout[-300:] + d.stderr
            q = p + ".dump.txt"
            want = (open(q, "rb").read() if os.path.exists(q) else gzip.open(q + ".gz", "rb").read()).decode().splitlines()
            assert sorted(l for l in lines if not l.startswith("#")) == want, f"{case['name']}/{f}: records differ"
            if os.access(REF_JF, os.X_OK):  # the unmodified jellyfish reads our file
                j = subprocess.run([REF_JF, "dump", "-c", os.path.join(out, f)], cwd=tmp_path, capture_output=True, text=True)
                assert j.returncode == 0 and sorted(j.stdout.splitlines()) == want
        else:
            want = open(p, "rb").read() if os.path.exists(p) else gzip.open(p + ".gz", "rb").read()
            assert open(tmp_path / out / f, "rb").read() == want, f"{case['name']}/{f} differs from the reference's file"

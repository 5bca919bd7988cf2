"""Host logic of the C++ front-end (kat_b200/bin/kat) WITHOUT a GPU: the same commands and the same
reference-written golden files as tests/test_gpu_cli.py, run against a test double of libkatgpu that
answers the C-ABI calls with the CPU oracle (tests/hostcheck/katgpu_double.c, interposed in front of the real
library with LD_PRELOAD).  What this covers is everything above the ABI — option handling, input
groups, FASTA/FASTQ readers, LOAD mode and hash dumps, the text writers, sect/cold/filter arithmetic.
What the GPU computes is covered by the `-m gpu` run of the same test functions."""
import os
import subprocess

import pytest

import test_gpu_cli as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOUBLE = os.path.join(ROOT, "tests", "hostcheck", "_build", "libkatgpu_double.so")


@pytest.fixture(scope="module", autouse=True)
def _double_in_front():
    if not os.access(T.KAT, os.X_OK):
        pytest.skip("kat_b200/bin/kat is not built (python __graft_entry__.py)")
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    old = os.environ.get("LD_PRELOAD")
    os.environ["LD_PRELOAD"] = DOUBLE + (":" + old if old else "")  # inherited by every `kat` the tests start
    os.environ["KAT_GPU_PARSE"] = "1"  # the double refuses raw FASTQ blocks, so the host route is what runs
    yield
    if old is None:
        del os.environ["LD_PRELOAD"]
    else:
        os.environ["LD_PRELOAD"] = old
    os.environ.pop("KAT_GPU_PARSE", None)


def test_the_double_is_what_runs(tmp_path):
    """The double answers: it reports one device unless told otherwise (the real library reports none on a machine without a GPU)."""
    g = subprocess.run([T.KAT, "hist", "-m5", "--gpus", "2", "-o", str(tmp_path / "h"), os.path.join(T.G.GOLDEN, "inputs", "edge1.fa")],
                       capture_output=True, text=True)
    assert g.returncode == 4 and "only 1 CUDA devices visible" in g.stderr


# fixtures and test functions of the GPU suite, collected here without the gpu mark
workdir = T.workdir
jf_workdir = T.jf_workdir

test_cli_sect_and_cold_output_files_byte_identical = T.test_cli_sect_and_cold_output_files_byte_identical
test_cli_load_mode_files_byte_identical = T.test_cli_load_mode_files_byte_identical
test_cli_comp_dump_hashes_and_symlink = T.test_cli_comp_dump_hashes_and_symlink
test_cli_filter_matches_reference = T.test_cli_filter_matches_reference
test_cli_filter_errors = T.test_cli_filter_errors
test_cli_wide_kmers_output_files_byte_identical = T.test_cli_wide_kmers_output_files_byte_identical
test_cli_wide_kmers_limits = T.test_cli_wide_kmers_limits
test_cli_wide_kmers_hash_files = T.test_cli_wide_kmers_hash_files
test_cli_hash_full_exit_code = T.test_cli_hash_full_exit_code
test_cli_missing_input = T.test_cli_missing_input
test_cli_dev_paths_are_pipes = T.test_cli_dev_paths_are_pipes


@pytest.mark.parametrize("case", T.G.cases(), ids=lambda c: c["name"])
def test_cli_output_files_byte_identical(case, workdir):
    T.test_cli_output_files_byte_identical(case, "direct", workdir)  # the counting path is a property of the real library


def test_cli_c1_ecoli_k27(tmp_path):
    T.test_cli_c1_ecoli_k27("auto", tmp_path)


def test_cli_dump_hash_is_a_jellyfish_file(jf_workdir):
    T.test_cli_dump_hash_is_a_jellyfish_file("direct", jf_workdir)


@pytest.mark.parametrize("case", [c for c in T.SECT_MANIFEST if c["name"] in ("sect_regions_gc_k17", "cold_asm_vs_reads_k21", "sect_fastq_records_k13")],
                         ids=lambda c: c["name"])
def test_lookup_windows_do_not_change_outputs(case, tmp_path, monkeypatch):
    """sect / cold / filter seq keep a window of records' k-mer positions resident (KAT_LOOKUP_WINDOW positions, 2^26 by
    default); with windows of a few hundred positions — many per file, records larger than a window included — the
    output files are still the reference's."""
    monkeypatch.setenv("KAT_LOOKUP_WINDOW", "700")
    T.test_cli_sect_and_cold_output_files_byte_identical(case, tmp_path)


@pytest.mark.parametrize("case", [c for c in T.FILTER_MANIFEST if c["name"].startswith("fseq")] + [c for c in T.WIDE_MANIFEST if c["args"][0] in ("sect", "cold", "filter")],
                         ids=lambda c: c["name"])
def test_lookup_windows_filter_and_wide(case, tmp_path, monkeypatch):
    monkeypatch.setenv("KAT_LOOKUP_WINDOW", "333")
    (T.test_cli_filter_matches_reference if case in T.FILTER_MANIFEST else T.test_cli_wide_kmers_output_files_byte_identical)(case, tmp_path)


@pytest.mark.parametrize("seed", range(12))
def test_sect_and_filter_seq_against_the_reference_binary_on_random_files(seed, tmp_path):
    """Differential run: random FASTA / FASTQ sequence files (ragged and multi-line records, empty records, IUPAC codes,
    missing final newline) through the unmodified reference `kat sect` / `kat filter seq` and through the front-end
    (SeqAn-style record reading, per-position lookups, statistics, record writers): same files, byte for byte."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kat")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/kat not built")
    import numpy as np
    from test_abi_and_host import _random_file
    fastq = seed % 2 == 0
    ext = ".fq" if fastq else ".fa"
    (tmp_path / ("s" + ext)).write_bytes(_random_file(np.random.default_rng(seed), fastq))
    (tmp_path / "h.fa").write_bytes(_random_file(np.random.default_rng(seed + 1000), False))
    no_preload = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
    for tool, args in (("sect", ["sect", "-m", "5", "-g", "-E", "-F", "-o", "OUT/x", "s" + ext, "h.fa"]),
                       ("fseq", ["filter", "seq", "-m", "5", "--stats", "-s", "-T", "0.5", "-o", "OUT/x", "--seq", "s" + ext, "h.fa"])):
        got = {}
        for who, exe, env in (("ref", ref, no_preload), ("our", T.KAT, os.environ)):
            out = who + "_" + tool
            for attempt in range(3):  # the reference reads an uninitialised terminal width when stdout is not a tty and may abort on it
                os.makedirs(tmp_path / out, exist_ok=True)
                r = subprocess.run([exe] + [a.replace("OUT", out) for a in args], cwd=tmp_path, capture_output=True, text=True, env=env)
                if r.returncode >= 0:
                    break
            got[who] = (r.returncode, {f: open(tmp_path / out / f, "rb").read() for f in sorted(os.listdir(tmp_path / out))})
        assert got["our"] == got["ref"], (seed, tool, got["ref"][0], got["our"][0])


@pytest.mark.parametrize("seed", list(range(18)) + [195])
def test_hist_gcp_comp_against_the_reference_binary_on_random_commands(seed, tmp_path):
    """Differential run over the option space: random FASTA / FASTQ inputs (one to three per command), random k (one- and
    two-word), -N/-O, histogram bounds, scales, bins, 5' trim, 2- and 3-input comp — the unmodified reference and the front-end
    write the same files byte for byte.  Seed 195 is the group in which the first file ends with an empty FASTQ record and the
    reference drops the second file (the parser's seq_len is not reset between files)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kat")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/kat not built")
    import numpy as np
    from test_abi_and_host import _random_file
    rng = np.random.default_rng(seed)
    names = []
    for i in range(3):
        fq = rng.random() < 0.5
        names.append("in%d.%s" % (i, "fq" if fq else "fa"))
        (tmp_path / names[-1]).write_bytes(_random_file(np.random.default_rng(seed * 10 + i), fq))
    k = int(rng.choice([3, 5, 7, 11, 17, 33, 40]))
    tool = ["hist", "gcp", "comp"][seed % 3]
    if tool == "hist":
        args = ["hist", "-m", str(k), "-H", "4096", "-o", "OUT/h"] + (["-N"] if rng.random() < 0.3 else [])
        if rng.random() < 0.6:
            args += ["-l", str(int(rng.integers(1, 5))), "-h", str(int(rng.integers(5, 40))), "-i", str(int(rng.integers(1, 4)))]
        if rng.random() < 0.4:
            args += ["--5ptrim", str(int(rng.integers(0, 6)))]
        args += names[:int(rng.integers(1, 4))]
    elif tool == "gcp":
        args = ["gcp", "-m", str(k), "-H", "4096", "-o", "OUT/g"] + (["-N"] if rng.random() < 0.3 else [])
        if rng.random() < 0.6:
            args += ["-x", str(float(rng.choice([0.1, 0.5, 2.0, 1.5]))), "-y", str(int(rng.integers(1, 30)))]
        args += names[:int(rng.integers(1, 4))]
    else:
        args = ["comp", "-m", str(k), "-H", "4096", "-I", "4096", "-J", "4096", "-o", "OUT/c"] + [f for f in ("-N", "-O", "-h", "-n") if rng.random() < 0.3]
        if rng.random() < 0.6:
            args += ["-i", str(int(rng.integers(2, 40))), "-j", str(int(rng.integers(2, 40)))]
        if rng.random() < 0.4:
            args += ["-x", str(float(rng.choice([0.5, 2.0]))), "-y", str(float(rng.choice([0.25, 1.5])))]
        args += names[:(3 if rng.random() < 0.3 else 2)]
    no_preload = {k_: v for k_, v in os.environ.items() if k_ != "LD_PRELOAD"}
    got = {}
    for who, exe, env in (("ref", ref, no_preload), ("our", T.KAT, os.environ)):
        for attempt in range(3):  # (the reference may abort on the uninitialised terminal width it reads)
            os.makedirs(tmp_path / who, exist_ok=True)
            r = subprocess.run([exe, *[a.replace("OUT", who) for a in args], "-t1"], cwd=tmp_path, capture_output=True, text=True, env=env)
            if r.returncode >= 0:
                break
        got[who] = (r.returncode, {f: open(tmp_path / who / f, "rb").read().replace(who.encode() + b"/", b"X/") for f in sorted(os.listdir(tmp_path / who))})
    assert got["our"] == got["ref"], (seed, " ".join(args), got["ref"][0], got["our"][0])


import fuzz_vs_reference as F


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("kind", F.KINDS)
def test_front_end_against_the_reference_binary_wider(kind, seed, tmp_path):
    """tests/fuzz_vs_reference.py, a few seeds of every kind on each CPU run (campaigns of hundreds of seeds per kind were run
    while the front-end was written): sect with all its switches, cold, filter kmer, filter seq single and paired (a second
    file that runs out first ends the run like SeqAn does: "Unexpected end of input.", exit code 5), gzip'd inputs with
    per-file trims and -t, jellyfish hashes in both directions, and file groups read by several parser streams (-t): a
    broken FASTQ record ends one stream and the others go on, a file in no supported format is fatal while the parser is
    built and later costs its stream the last 4 KiB buffer of the previous file."""
    if not os.access(F.REF, os.X_OK):
        pytest.skip("oracle/_ref/kat not built")
    try:
        F.case(kind, seed, tmp_path)
    except F.ReferenceHung as e:
        pytest.skip("the reference did not end: " + str(e))


@pytest.mark.parametrize("seed", range(5, 9))
@pytest.mark.parametrize("kind", [k for k in F.KINDS if k not in ("load", "jfcount")])
def test_front_end_as_a_job_against_the_reference_binary(kind, seed, tmp_path):
    """The same differential cases with the front-end running as a job of 2 devices (of the test double): what a job takes
    (tests/fuzz_vs_reference.py::on_a_job) gives the reference's files."""
    if not os.access(F.REF, os.X_OK):
        pytest.skip("oracle/_ref/kat not built")
    try:
        if F.case(kind, seed, tmp_path, gpus=2) is False:
            pytest.skip("not a command for a job (hash files, k > 32 or non-canonical comp)")
    except F.ReferenceHung as e:
        pytest.skip("the reference did not end: " + str(e))


# ---- `kat --gpus N` on the CPU: N contexts of the double joined into one job (katgpu_double.c) ----
def _job_cases():
    out = []
    for c in T.G.cases():
        out.append(("golden", c))
    for c in T.SECT_MANIFEST:
        out.append(("sect", c))
    for c in T.FILTER_MANIFEST:
        out.append(("filter", c))
    # what a job does not do, and the front-end says so: dump or load jellyfish hashes (one GPU's table), count two-word
    # k-mers, compare inputs that were not all counted canonically (the shared-k-mer walk is local to the owner of a k-mer)
    def on_a_job(c):
        a = c["args"]
        return not ("-d" in a or "--dump_hash" in a or any(x.endswith(".jf") for x in a) or (a[0] == "comp" and {"-N", "-O", "-P"} & set(a)))
    return [(kind, c) for kind, c in out if on_a_job(c)]


@pytest.mark.parametrize("gpus", [2, 3])
@pytest.mark.parametrize("kind,case", _job_cases(), ids=lambda x: x["name"] if isinstance(x, dict) else x)
def test_cli_on_a_job_of_several_devices(kind, case, gpus, workdir, tmp_path, monkeypatch):
    """Every golden command again as `kat --gpus 2` and `--gpus 3` (KATGPU_GPUS): the front-end creates one context per
    device, joins them, cuts every parsed base stream into per-GPU shares of k-mer start positions with k-1 bases of
    overlap (kat_front.hpp::feed_all), sizes the per-GPU tables, calls every walk and lookup collectively from one host
    thread per GPU and merges per-rank parts (filter kmer) — and writes the reference's files byte for byte.  On the GPU
    box the same is checked against the real library by tests/test_gpu_multi.py where two GPUs are present."""
    monkeypatch.setenv("KATGPU_DOUBLE_DEVICES", str(gpus))
    monkeypatch.setenv("KATGPU_GPUS", str(gpus))
    if kind == "golden":
        T.test_cli_output_files_byte_identical(case, "direct", workdir)
    elif kind == "sect":
        T.test_cli_sect_and_cold_output_files_byte_identical(case, tmp_path)
    else:
        T.test_cli_filter_matches_reference(case, tmp_path)


@pytest.mark.parametrize("gpus", [1, 2])
@pytest.mark.parametrize("kind,case", [kc for kc in _job_cases() if any(a.endswith((".fq", ".fastq")) for a in kc[1]["args"])],
                         ids=lambda x: x["name"] if isinstance(x, dict) else x)
def test_cli_raw_fastq_route(kind, case, gpus, workdir, tmp_path, monkeypatch):
    """The front-end's device-parsing route on the CPU: with KATGPU_DOUBLE_FASTQ=1 the double takes raw FASTQ blocks under the
    real library's rules (all or nothing per call), so stream_fastq_raw's cutting of a file into 4 KiB blocks of whole records
    and per-GPU shares, the per-file trims, and the fall back to the host route — from the start of the input when a later
    block is refused — run for every golden command with a FASTQ input, on one device and on a job of two."""
    monkeypatch.setenv("KATGPU_DOUBLE_FASTQ", "1")
    monkeypatch.setenv("KAT_RAW_BLOCK", "4096")
    monkeypatch.setenv("KATGPU_DOUBLE_DEVICES", str(gpus))
    monkeypatch.setenv("KATGPU_GPUS", str(gpus))
    if kind == "golden":
        T.test_cli_output_files_byte_identical(case, "direct", workdir)
    elif kind == "sect":
        T.test_cli_sect_and_cold_output_files_byte_identical(case, tmp_path)
    else:
        T.test_cli_filter_matches_reference(case, tmp_path)


@pytest.mark.parametrize("seed", range(9, 12))
@pytest.mark.parametrize("kind", ["gz", "group", "compgroups", "cold"])
def test_front_end_raw_fastq_route_against_the_reference_binary(kind, seed, tmp_path, monkeypatch):
    """... and the differential cases with FASTQ inputs through the same route."""
    if not os.access(F.REF, os.X_OK):
        pytest.skip("oracle/_ref/kat not built")
    monkeypatch.setenv("KATGPU_DOUBLE_FASTQ", "1")
    monkeypatch.setenv("KAT_RAW_BLOCK", "4096")
    try:
        F.case(kind, seed, tmp_path)
    except F.ReferenceHung as e:
        pytest.skip("the reference did not end: " + str(e))


def test_usage_errors_end_like_the_reference(tmp_path):
    """Exit codes (and the one-line messages) of commands that do no work: the reference's main drops what a tool's main returns,
    so a tool that only printed its help ends with 0 (src/kat.cc:259-284); a fourth input of comp is a command-line error (1);
    a trim that is no number is boost::bad_lexical_cast (4); a mer length of 0 or a hash size of 1 give a hash matrix without
    columns / rows (5)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "kat")
    if not os.access(ref, os.X_OK):
        pytest.skip("oracle/_ref/kat not built")
    fa, fb = os.path.join(T.G.GOLDEN, "inputs", "edge1.fa"), os.path.join(T.G.GOLDEN, "inputs", "synth_c.fa")
    no_preload = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
    for args, needle in ((["hist", "--help"], ""), (["filter"], ""), (["filter", "kmer"], ""), (["filter", "seq"], ""), (["--version"], ""),
                         (["bogus"], ""), (["comp", "-m5", "a", "b", "c", "d"], "too many positional options"),
                         (["comp", fa], ""), (["sect", fa], ""), (["cold", fa], ""),
                         (["hist", "-m5", "-o", "x/h", "--5ptrim", "x", fa], ""), (["hist", "-m5", "-o", "x/h", "--5ptrim", "1,2", fa], "Inconsistent number of inputs"),
                         (["hist", "-m", "0", "-o", "x/h", fa], "Invalid matrix size 27x0"), (["hist", "-m", "0", "-H", "1", "-o", "x/h", fa], "Invalid matrix size 0x0"),
                         (["comp", "-m", "0", "-o", "x/c", fa, fb], "Invalid matrix size 27x0"), (["gcp", "-m5", "-H", "1", "-o", "x/g", fa], "Invalid matrix size 0x10"),
                         (["gcp", "-m5", "-y", "65536", "-o", "x/g", fa], "for option '--cvg_bins' is invalid"),
                         (["sect", "-m5", "-G", "5000000000", "-o", "x/s", fa, fb], "for option '--max_repeat' is invalid"),
                         (["hist", "-m5", "-t", "70000", "-o", "x/h", fa], "for option '--threads' is invalid"),
                         # path-typed options end at the first blank for boost: a file name with a blank is a command-line error
                         (["hist", "-m5", "-o", "x/h", fa, "we ird.fa"], "('we ird.fa') for option '--inputs' is invalid"),
                         (["sect", "-m5", "-o", "x/s", "we ird.fa", fa], "('we ird.fa') for option '--seq_file' is invalid"),
                         (["cold", "-m5", "-o", "x/c", fa, "we ird.fa"], "('we ird.fa') for option '--reads_files' is invalid"),
                         (["gcp", "-m5", "-o", "x y/g", fa], "('x y/g') for option '--output_prefix' is invalid"),
                         (["comp", "-m5", "-o", "x y/c", fa, fb], "")):  # (comp takes its prefix, like its inputs, as a string)
        got = []
        for exe, env in ((ref, no_preload), (T.KAT, os.environ)):
            for attempt in range(5):  # (the reference may abort on the uninitialised terminal width it reads)
                r = subprocess.run([exe] + args, cwd=tmp_path, capture_output=True, text=True, env=env, timeout=120)
                if r.returncode >= 0:
                    break
            assert needle in r.stderr, (exe, args, r.stderr[-300:])
            got.append(r.returncode)
        assert got[0] == got[1], (args, got)

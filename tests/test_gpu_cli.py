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


@pytest.mark.parametrize("path", ["direct", "partition"])
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


def test_cli_hash_full_exit_code(workdir):
    """--disable_hash_grow + a table that is too small: the reference throws
    std::runtime_error("Hash full") -> "Error: Hash full", exit code 5 (src/kat.cc:292-294)."""
    r = subprocess.run([KAT, "comp", "-m21", "-g", "-H", "1000", "-I", "1000", "-o", "hf/c", "inputs/synth_a.fq", "inputs/synth_c.fa"],
                       cwd=workdir, capture_output=True, text=True, timeout=300)
    assert r.returncode == 5 and "Hash full" in r.stderr


def test_cli_missing_input(workdir):
    r = subprocess.run([KAT, "hist", "-m21", "inputs/nope.fa"], cwd=workdir, capture_output=True, text=True)
    assert r.returncode == 4 and "Could not find input file" in r.stderr

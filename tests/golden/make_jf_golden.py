#!/usr/bin/env python
"""Golden fixtures for the jellyfish-hash boundary (LOAD mode and -d dumps):
hash files written by the UNMODIFIED reference binaries (tests/golden/jf/*.jf), their
record listing by the reference's own `jellyfish dump -c`, and the outputs of the reference
`kat hist|gcp|comp` when those hashes are its INPUT (tests/golden/jf/expected/...).

Run from the repo root where /root/reference exists, after make_golden.py:
    make -C oracle ref && python tests/golden/make_jf_golden.py"""
import gzip
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
JF = os.path.join(HERE, "jf")


def run(cmd, **kw):
    r = subprocess.run(cmd, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise SystemExit(f"failed: {cmd}\n{r.stdout[-800:]}")
    return r.stdout


# hashes to create: name -> command producing jf/<name>.jf (OUT is replaced by a scratch prefix)
HASHES = [
    # written by the reference KAT's own -d (InputHandler::dump): canonical 13-mers of long.fa
    ("long_k13_kat", ["kat", "hist", "-m13", "-d", "-o", "OUT/h", "inputs/long.fa"], "OUT/h-hash.jf13"),
    # written by the bundled jellyfish `count` (non-canonical, 2-byte counters)
    ("polyA_edge_k5_jf", ["kat_jellyfish", "count", "-m", "5", "-s", "4096", "--out-counter-len", "2", "-o", "OUT/x.jf",
                          "inputs/polyA.fa", "inputs/edge1.fa"], "OUT/x.jf"),
    # canonical 21-mers counted by jellyfish from the assembly-like FASTA
    ("synthc_k21_jf", ["kat_jellyfish", "count", "-C", "-m", "21", "-s", "50000", "-o", "OUT/y.jf", "inputs/synth_c.fa"], "OUT/y.jf"),
]
# the reference's OWN hash fixture (tests/data/ecoli.header.jf27, 22 KB, written by jellyfish with a
# random hash matrix): its tests expect 1889 records and four named counts (tests/check_jellyfish.cc:38-116)
REF_FIXTURE = ("/root/reference/tests/data/ecoli.header.jf27", "ref_ecoli_header_k27")

# reference KAT runs that take a hash as input (LOAD mode); paths relative to tests/golden
CASES = [
    ("load_hist_ref_fixture_k27", ["hist", "-o", "OUT/h", "jf/ref_ecoli_header_k27.jf"]),
    ("load_hist_k13", ["hist", "-o", "OUT/h", "jf/long_k13_kat.jf"]),
    ("load_hist_k5_lowhigh", ["hist", "-l", "2", "-h", "40000", "-i", "7", "-o", "OUT/h", "jf/polyA_edge_k5_jf.jf"]),
    ("load_gcp_k21", ["gcp", "-y", "30", "-o", "OUT/g", "jf/synthc_k21_jf.jf"]),
    ("load_comp_hash_vs_reads_k21", ["comp", "-m21", "-i", "40", "-j", "50", "-h", "-o", "OUT/c", "jf/synthc_k21_jf.jf", "inputs/synth_a.fq"]),
    ("load_comp_two_hashes_k13", ["comp", "-i", "30", "-j", "30", "-o", "OUT/c", "jf/long_k13_kat.jf", "jf/long_k13_kat.jf"]),
]


def main():
    shutil.rmtree(JF, ignore_errors=True)
    os.makedirs(os.path.join(JF, "expected"))
    scratch = os.path.join(JF, "_scratch")
    manifest = {"hashes": [], "cases": []}
    for name, cmd, produced in HASHES:
        os.makedirs(scratch, exist_ok=True)
        args = [os.path.join(REF, cmd[0])] + [a.replace("OUT", "jf/_scratch") for a in cmd[1:]]
        run(args)
        dst = os.path.join(JF, name + ".jf")
        shutil.move(os.path.join(HERE, produced.replace("OUT", "jf/_scratch")), dst)
        shutil.rmtree(scratch)
        with gzip.GzipFile(os.path.join(JF, name + ".dump.txt.gz"), "wb", mtime=0) as f:
            f.write(run([os.path.join(REF, "kat_jellyfish"), "dump", "-c", "jf/" + name + ".jf"]).encode())
        manifest["hashes"].append({"name": name, "cmd": cmd})
        print(name, os.path.getsize(dst), "bytes")
    shutil.copy(REF_FIXTURE[0], os.path.join(JF, REF_FIXTURE[1] + ".jf"))
    with gzip.GzipFile(os.path.join(JF, REF_FIXTURE[1] + ".dump.txt.gz"), "wb", mtime=0) as f:
        f.write(run([os.path.join(REF, "kat_jellyfish"), "dump", "-c", "jf/" + REF_FIXTURE[1] + ".jf"]).encode())
    manifest["hashes"].append({"name": REF_FIXTURE[1], "cmd": ["copy of the reference's tests/data/ecoli.header.jf27"]})
    for name, args in CASES:
        out = os.path.join("jf", "expected", name)
        os.makedirs(os.path.join(HERE, out))
        run([os.path.join(REF, "kat")] + [a.replace("OUT", out) for a in args])
        files = sorted(os.listdir(os.path.join(HERE, out)))
        for f in files:  # keep the repo small
            p = os.path.join(HERE, out, f)
            if os.path.getsize(p) > 100_000:
                with open(p, "rb") as src, gzip.GzipFile(p + ".gz", "wb", mtime=0) as dst:
                    dst.write(src.read())
                os.unlink(p)
        manifest["cases"].append({"name": name, "args": args, "files": files})
        print(name, files)
    with open(os.path.join(JF, "cases.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()

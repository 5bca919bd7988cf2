#!/usr/bin/env python
"""Generate the golden fixtures: small deterministic inputs (tests/golden/inputs)
and the outputs of the UNMODIFIED reference binary on them
(tests/golden/expected/<case>/...), plus cases.json describing each command.

Run from the repo root in the container that has /root/reference:
    make -C oracle ref && python tests/golden/make_golden.py
The reference prints input paths into its headers, so every command runs with
cwd = tests/golden and relative paths; the product is run the same way."""
import gzip
import json
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_KAT = os.path.join(ROOT, "oracle", "_ref", "kat")
INP = os.path.join(HERE, "inputs")
EXP = os.path.join(HERE, "expected")
ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def w(name, data):
    with open(os.path.join(INP, name), "wb") as f:
        f.write(data)


def reads_from(genome, n, L, rng, sub=0.01, n_rate=0.002):
    pos = rng.integers(0, len(genome) - L + 1, n)
    r = genome[pos[:, None] + np.arange(L)[None, :]].copy()
    rev = rng.random(n) < 0.5
    r[rev] = 3 - r[rev][:, ::-1]
    m = rng.random(r.shape) < sub
    r[m] = (r[m] + rng.integers(1, 4, int(m.sum()))) & 3
    s = ALPHA[r]
    s[rng.random(r.shape) < n_rate] = ord("N")
    return s


def fastq(reads, prefix):
    out = bytearray()
    for i, r in enumerate(reads):
        out += b"@%s%d\n" % (prefix, i) + r.tobytes() + b"\n+\n" + b"I" * len(r) + b"\n"
    return bytes(out)


def fasta_wrapped(name, seq, width=70):
    s = seq.tobytes()
    return b">" + name + b"\n" + b"\n".join(s[i:i + width] for i in range(0, len(s), width)) + b"\n"


def make_inputs():
    shutil.rmtree(INP, ignore_errors=True)
    os.makedirs(INP)
    # hand-made edge cases
    w("edge1.fa", b">r1\nACGTAcgtNACGTA\nCGT\n>r2 short\nACG\n>r3 iupac\nAAAAARAAAAA\n>r4\nTTTTTTT\n")
    w("edge2.fq", b"@r1\nACGTACGTAC\n+\n@IIIIIIIII\n@r2\nACGTA\nCGTAC\n+r2\nIIIII\nIIIII\n@r3\nGGGGGNGGGGG\n+\nIIIIIIIIIII\n")
    w("crlf.fa", b">c\r\nACGTACG\r\nTACGT\r\n")
    w("nonl.fq", b"@r1\nACGTACGTAC\n+\nIIIIIIIIII\n@r2\nTTTTTGGGGG\n+\nIIIIIIIIII")
    w("polyA.fa", b"".join(b">p%d\n" % i + b"A" * 104 + b"\n" for i in range(300)))
    w("empty.fa", b"")
    rng = np.random.default_rng(20260925)
    g = rng.integers(0, 4, 20000).astype(np.uint8)
    w("synth_a.fq", fastq(reads_from(g, 2000, 100, rng), b"a"))
    w("synth_b.fq", fastq(reads_from(g, 1500, 100, rng, sub=0.02), b"b"))
    with gzip.GzipFile(os.path.join(INP, "synth_a.fq.gz"), "wb", mtime=0) as f:
        f.write(open(os.path.join(INP, "synth_a.fq"), "rb").read())
    asm = fasta_wrapped(b"ctg1", ALPHA[g[:9000]]) + fasta_wrapped(b"ctg2 second", ALPHA[g[9500:15000]]) + \
        b">ctg3\n" + ALPHA[g[15000:]].tobytes() + b"\n"
    w("synth_c.fa", asm)
    # one long record, partly on one line, partly wrapped: crosses several 4 KiB parser buffers
    lg = ALPHA[rng.integers(0, 4, 12000)]
    w("long.fa", b">long\n" + lg[:6000].tobytes() + b"\n" + fasta_wrapped(b"x", lg[6000:])[3:])


CASES = [
    ("hist_edge1_nc", ["hist", "-m5", "-N", "-o", "OUT/h", "inputs/edge1.fa"]),
    ("hist_edge1_canon", ["hist", "-m5", "-o", "OUT/h", "inputs/edge1.fa"]),
    ("hist_edge1_lhi", ["hist", "-m5", "-N", "-l", "2", "-h", "6", "-i", "2", "-o", "OUT/h", "inputs/edge1.fa"]),
    ("hist_edge2_nc", ["hist", "-m5", "-N", "-o", "OUT/h", "inputs/edge2.fq"]),
    ("hist_edge2_trim2", ["hist", "-m5", "-N", "--5ptrim", "2", "-o", "OUT/h", "inputs/edge2.fq"]),
    ("hist_crlf", ["hist", "-m5", "-N", "-o", "OUT/h", "inputs/crlf.fa"]),
    ("hist_nonl", ["hist", "-m5", "-N", "-o", "OUT/h", "inputs/nonl.fq"]),
    ("hist_polyA", ["hist", "-m5", "-o", "OUT/h", "inputs/polyA.fa"]),
    ("hist_two_files_k27", ["hist", "-m27", "-t2", "-o", "OUT/h", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("hist_trim_per_file_k17", ["hist", "-m17", "--5ptrim", "3,5", "-o", "OUT/h", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("hist_fasta_trim_quirk_k21", ["hist", "-m21", "-N", "--5ptrim", "7", "-o", "OUT/h", "inputs/long.fa"]),
    ("hist_asm_k31", ["hist", "-m31", "-o", "OUT/h", "inputs/synth_c.fa"]),
    ("hist_k32_gz", ["hist", "-m32", "-H", "100000", "-o", "OUT/h", "inputs/synth_a.fq.gz"]),
    ("hist_grow_small_H", ["hist", "-m21", "-H", "1000", "-o", "OUT/h", "inputs/synth_a.fq", "inputs/synth_c.fa"]),
    ("gcp_edge1", ["gcp", "-m5", "-x", "0.5", "-y", "3", "-o", "OUT/g", "inputs/edge1.fa"]),
    ("gcp_k17", ["gcp", "-m17", "-o", "OUT/g", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("gcp_k27_scaled_gz", ["gcp", "-m27", "-x", "0.25", "-y", "50", "-o", "OUT/g", "inputs/synth_a.fq.gz"]),
    ("comp_k13_hists", ["comp", "-m13", "-h", "-i", "120", "-j", "40", "-o", "OUT/c", "inputs/synth_a.fq", "inputs/synth_c.fa"]),
    ("comp_k21_density_scaled", ["comp", "-m21", "-n", "-i", "50", "-j", "60", "-x", "0.5", "-y", "2", "-o", "OUT/c",
                                 "inputs/synth_a.fq inputs/synth_b.fq", "inputs/synth_c.fa"]),
    ("comp_k17_three_inputs", ["comp", "-m17", "-i", "80", "-j", "70", "-o", "OUT/c", "inputs/synth_a.fq", "inputs/synth_b.fq",
                               "inputs/synth_c.fa"]),
    ("comp_k21_mixed_canonical", ["comp", "-m21", "-N", "-i", "60", "-j", "60", "-h", "-o", "OUT/c", "inputs/synth_a.fq",
                                  "inputs/synth_c.fa"]),
    ("comp_k15_glob_default_bins", ["comp", "-m15", "-o", "OUT/c", "inputs/synth_?.fq", "inputs/synth_c.fa"]),
]


def main():
    if not os.access(REF_KAT, os.X_OK):
        raise SystemExit("oracle/_ref/kat missing: run `make -C oracle ref` where /root/reference exists")
    make_inputs()
    shutil.rmtree(EXP, ignore_errors=True)
    manifest = []
    for name, args in CASES:
        out_dir = os.path.join(EXP, name)
        os.makedirs(out_dir)
        a = [x.replace("OUT", os.path.join("expected", name)) for x in args]
        r = subprocess.run([REF_KAT] + a, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit(f"reference failed on {name}: {r.stdout[-800:]}")
        files = sorted(os.listdir(out_dir))
        for f in files:  # keep the repo small: big matrices are stored gzipped
            p = os.path.join(out_dir, f)
            if os.path.getsize(p) > 200_000:
                with open(p, "rb") as src, gzip.GzipFile(p + ".gz", "wb", mtime=0) as dst:
                    dst.write(src.read())
                os.unlink(p)
        manifest.append({"name": name, "args": args, "files": files})
        print(name, files)
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()

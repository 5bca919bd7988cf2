#!/usr/bin/env python
"""Golden fixtures for `kat filter seq` and `kat filter kmer`: what the UNMODIFIED
reference binary writes for a handful of commands over tests/golden/inputs, stored under tests/golden/filter/.
filter seq: the output files themselves.  filter kmer: the hashes it writes hold the reference's own random hash
matrix, so what is stored is their content — `kat_jellyfish dump -c`, sorted — as <file>.dump.txt.
Run from the repo root where /root/reference exists:  make -C oracle ref && python tests/golden/make_filter_golden.py
Commands run with cwd = tests/golden and relative paths, as make_golden.py does."""
import gzip
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_KAT = os.path.join(ROOT, "oracle", "_ref", "kat")
REF_JF = os.path.join(ROOT, "oracle", "_ref", "kat_jellyfish")
OUT = os.path.join(HERE, "filter")

CASES = [
    ("fseq_fastq_stats_separate_k13", ["filter", "seq", "-m13", "-H", "1000000", "--stats", "-s", "-o", "OUT/f", "--seq", "inputs/synth_a.fq", "inputs/synth_c.fa"]),
    ("fseq_fasta_invert_wrap_k21", ["filter", "seq", "-m21", "-i", "-T", "0.5", "-s", "--stats", "-o", "OUT/f", "--seq", "inputs/synth_c.fa", "inputs/synth_b.fq"]),
    ("fseq_paired_k17", ["filter", "seq", "-m17", "-T", "0.3", "-s", "--stats", "-o", "OUT/f", "--seq", "inputs/pair_1.fq", "--seq2", "inputs/pair_2.fq", "inputs/synth_c.fa"]),
    ("fseq_multiline_short_nc_k5", ["filter", "seq", "-m5", "-N", "--stats", "-s", "-o", "OUT/f", "--seq", "inputs/edge2.fq", "inputs/edge1.fa"]),
    ("fseq_iupac_fasta_k5", ["filter", "seq", "-m5", "--stats", "-T", "0.75", "-s", "-o", "OUT/f", "--seq", "inputs/edge1.fa", "inputs/edge2.fq"]),
    ("fseq_long_fasta_two_inputs_k27", ["filter", "seq", "-m27", "--stats", "-o", "OUT/f", "--seq", "inputs/long.fa", "inputs/synth_c.fa", "inputs/long.fa"]),
    ("fseq_loaded_hash_k21", ["filter", "seq", "--stats", "-s", "-T", "0.2", "-o", "OUT/f", "--seq", "inputs/synth_a.fq", "jf/synthc_k21_jf.jf"]),
    ("fkmer_bounds_separate_k15", ["filter", "kmer", "-m15", "-H", "100000", "-c", "2", "-d", "5", "-g", "5", "-h", "9", "-s", "-o", "OUT/k", "inputs/synth_a.fq"]),
    ("fkmer_invert_nc_two_files_k21", ["filter", "kmer", "-m21", "-N", "-i", "-c", "2", "-d", "3", "-g", "8", "-h", "12", "-o", "OUT/k", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("fkmer_defaults_k17", ["filter", "kmer", "-m17", "-o", "OUT/k", "inputs/synth_c.fa"]),
    ("fkmer_separate_ignores_invert_k13", ["filter", "kmer", "-m13", "-s", "-i", "-c", "3", "-o", "OUT/k", "inputs/long.fa"]),
    ("fkmer_loaded_hash_k21", ["filter", "kmer", "-c", "1", "-d", "1", "-g", "0", "-h", "10", "-s", "-o", "OUT/k", "jf/synthc_k21_jf.jf"]),
]


def make_pair_inputs():
    """inputs/pair_1.fq / pair_2.fq: 1500 mates each (the first 1500 records of synth_a.fq, and synth_b.fq)."""
    a = open(os.path.join(HERE, "inputs", "synth_a.fq")).read().splitlines(keepends=True)
    with open(os.path.join(HERE, "inputs", "pair_1.fq"), "w") as f:
        f.writelines(a[: 4 * 1500])
    shutil.copy(os.path.join(HERE, "inputs", "synth_b.fq"), os.path.join(HERE, "inputs", "pair_2.fq"))


def main():
    make_pair_inputs()
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    manifest = []
    for name, args in CASES:
        d = os.path.join(OUT, name)
        os.makedirs(d)
        a = [x.replace("OUT", os.path.join("filter", name)) for x in args]
        r = subprocess.run([REF_KAT] + a, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit(f"reference failed on {name}: {r.stdout[-800:]}")
        files = sorted(os.listdir(d))
        for f in files:
            p = os.path.join(d, f)
            if ".jf" in f:  # a hash: keep its records as sorted text
                j = subprocess.run([REF_JF, "dump", "-c", p], stdout=subprocess.PIPE, text=True, check=True)
                os.unlink(p)
                p = p + ".dump.txt"
                with open(p, "w") as out:
                    out.writelines(sorted(j.stdout.splitlines(keepends=True)))
            if os.path.getsize(p) > 50_000:
                with open(p, "rb") as src, gzip.GzipFile(p + ".gz", "wb", mtime=0) as dst:
                    dst.write(src.read())
                os.unlink(p)
        summary = [l for l in r.stdout.splitlines() if l.startswith(("K-mers ", "Found "))]
        manifest.append({"name": name, "args": args, "files": files, "summary": summary})
        print(name, files, summary)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()

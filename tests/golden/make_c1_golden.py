#!/usr/bin/env python
"""BASELINE.json configs[0] (SURVEY.md §8 "C1") as a golden fixture: the reference's own test
genome tests/data/EcoliK12.fasta (and its read sample ecoli_r1.1K.fastq), gzip-compressed, and the
files the UNMODIFIED reference binary writes for
    kat hist -m27 -t1 c1/EcoliK12.fasta.gz
    kat comp -m13 -t1 c1/ecoli_r1.1K.fastq.gz c1/EcoliK12.fasta.gz        (SURVEY.md §8(c) probe)
Run from the repo root in the container that has /root/reference:
    make -C oracle ref && python tests/golden/make_c1_golden.py
Commands run with cwd = tests/golden and relative paths (paths appear in the headers)."""
import gzip
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_KAT = os.path.join(ROOT, "oracle", "_ref", "kat")
SRC = "/root/reference/tests/data"
C1 = os.path.join(HERE, "c1")


def main():
    os.makedirs(C1, exist_ok=True)
    for name in ("EcoliK12.fasta", "ecoli_r1.1K.fastq"):
        with open(os.path.join(SRC, name), "rb") as f, gzip.GzipFile(os.path.join(C1, name + ".gz"), "wb", 9, mtime=0) as g:
            shutil.copyfileobj(f, g)
    out = os.path.join(C1, "out")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    subprocess.check_call([REF_KAT, "hist", "-m27", "-t1", "-o", "c1/out/ecoli.hist", "c1/EcoliK12.fasta.gz"], cwd=HERE)
    subprocess.check_call([REF_KAT, "comp", "-m13", "-t1", "-o", "c1/out/r1_vs_k12", "c1/ecoli_r1.1K.fastq.gz", "c1/EcoliK12.fasta.gz"], cwd=HERE)
    keep = {"ecoli.hist", "r1_vs_k12-main.mx", "r1_vs_k12.stats"}
    for f in sorted(os.listdir(out)):
        p = os.path.join(out, f)
        if f not in keep:
            os.unlink(p)
            continue
        with open(p, "rb") as fi, gzip.GzipFile(p + ".gz", "wb", 9, mtime=0) as g:
            shutil.copyfileobj(fi, g)
        os.unlink(p)
    print(sorted(os.listdir(out)))


if __name__ == "__main__":
    main()

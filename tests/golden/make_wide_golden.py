#!/usr/bin/env python
"""Golden fixtures for mer lengths above 32 (two-word keys): the files the UNMODIFIED reference
binary writes for hist / gcp / comp / sect / cold / filter seq at k = 33..64 over tests/golden/inputs, stored under
tests/golden/wide/.  Run from the repo root where /root/reference exists:
  make -C oracle ref && python tests/golden/make_wide_golden.py
Commands run with cwd = tests/golden and relative paths, as make_golden.py does."""
import gzip
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_KAT = os.path.join(ROOT, "oracle", "_ref", "kat")
OUT = os.path.join(HERE, "wide")

CASES = [
    ("hist_k33_two_files", ["hist", "-m33", "-o", "OUT/h", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("hist_k64_non_canonical", ["hist", "-m64", "-N", "-o", "OUT/h", "inputs/synth_a.fq"]),
    ("hist_k48_polyA_bounds", ["hist", "-m48", "-l", "2", "-h", "60", "-i", "3", "-o", "OUT/h", "inputs/polyA.fa", "inputs/long.fa"]),
    ("hist_k63_gz_trim", ["hist", "-m63", "--5ptrim", "3", "-o", "OUT/h", "inputs/synth_a.fq.gz"]),
    ("hist_k40_edges_crlf", ["hist", "-m40", "-N", "-o", "OUT/h", "inputs/crlf.fa", "inputs/edge1.fa", "inputs/edge2.fq", "inputs/nonl.fq", "inputs/long.fa"]),
    ("gcp_k48", ["gcp", "-m48", "-o", "OUT/g", "inputs/synth_a.fq"]),
    ("gcp_k40_scaled", ["gcp", "-m40", "-x", "0.5", "-y", "50", "-o", "OUT/g", "inputs/synth_b.fq", "inputs/synth_c.fa"]),
    ("comp_k41_hists", ["comp", "-m41", "-h", "-i", "100", "-j", "50", "-o", "OUT/c", "inputs/synth_a.fq", "inputs/synth_c.fa"]),
    ("comp_k35_three_inputs", ["comp", "-m35", "-i", "80", "-j", "70", "-o", "OUT/c", "inputs/synth_a.fq", "inputs/synth_b.fq", "inputs/synth_c.fa"]),
    ("comp_k64_mixed_canonical", ["comp", "-m64", "-N", "-i", "60", "-j", "60", "-h", "-o", "OUT/c", "inputs/synth_a.fq", "inputs/synth_c.fa"]),
    ("sect_k37_regions", ["sect", "-m37", "-E", "-F", "-g", "-M", "3", "-G", "40", "-o", "OUT/s", "inputs/synth_c.fa", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("sect_k33_fastq_nc", ["sect", "-m33", "-N", "-o", "OUT/s", "inputs/synth_b.fq", "inputs/synth_a.fq"]),
    ("cold_k45", ["cold", "-m45", "-o", "OUT/c", "inputs/synth_c.fa", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("fseq_k40_stats", ["filter", "seq", "-m40", "--stats", "-s", "-T", "0.4", "-o", "OUT/f", "--seq", "inputs/synth_a.fq", "inputs/synth_c.fa"]),
]


REF_JF = os.path.join(ROOT, "oracle", "_ref", "kat_jellyfish")
# jellyfish hashes with two-word keys either side of the path: a hash written by the reference as INPUT (LOAD mode),
# and the records of the hashes the reference dumps / filters (kat_jellyfish dump -c, sorted) to compare ours with
HASH_CASES = [
    ("load_hist_k41", ["hist", "-o", "OUT/h", "wide/ref_small_k41.jf"]),
    ("load_gcp_k41", ["gcp", "-o", "OUT/g", "wide/ref_small_k41.jf"]),
    ("load_comp_hash_vs_reads_k41", ["comp", "-m41", "-i", "40", "-j", "50", "-o", "OUT/c", "wide/ref_small_k41.jf", "inputs/synth_a.fq"]),
    ("dump_hist_k45_nc", ["hist", "-m45", "-N", "-d", "-o", "OUT/h", "inputs/wide_small.fq"]),
    ("fkmer_k33_separate", ["filter", "kmer", "-m33", "-c", "2", "-d", "4", "-g", "10", "-h", "20", "-s", "-o", "OUT/k", "inputs/wide_small.fq"]),
]


def make_hash_cases():
    a = open(os.path.join(HERE, "inputs", "synth_a.fq")).read().splitlines(keepends=True)
    with open(os.path.join(HERE, "inputs", "wide_small.fq"), "w") as f:
        f.writelines(a[: 4 * 60])
    subprocess.run([REF_KAT, "hist", "-m41", "-d", "-o", "wide/tmp_h", "inputs/wide_small.fq"], cwd=HERE, check=True, stdout=subprocess.DEVNULL)
    os.replace(os.path.join(OUT, "tmp_h-hash.jf41"), os.path.join(OUT, "ref_small_k41.jf"))
    os.unlink(os.path.join(OUT, "tmp_h"))
    manifest = []
    for name, args in HASH_CASES:
        d = os.path.join(OUT, name)
        os.makedirs(d)
        a = [x.replace("OUT", os.path.join("wide", name)) for x in args]
        r = subprocess.run([REF_KAT] + a, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit(f"reference failed on {name}: {r.stdout[-800:]}")
        files = sorted(os.listdir(d))
        for f in files:
            p = os.path.join(d, f)
            if ".jf" in f:
                j = subprocess.run([REF_JF, "dump", "-c", p], stdout=subprocess.PIPE, text=True, check=True)
                os.unlink(p)
                p = p + ".dump.txt"
                with open(p, "w") as out:
                    out.writelines(sorted(j.stdout.splitlines(keepends=True)))
            if os.path.getsize(p) > 30_000:
                with open(p, "rb") as src, gzip.GzipFile(p + ".gz", "wb", mtime=0) as dst:
                    dst.write(src.read())
                os.unlink(p)
        manifest.append({"name": name, "args": args, "files": files})
        print(name, files)
    with open(os.path.join(OUT, "hash_cases.json"), "w") as f:
        json.dump(manifest, f, indent=1)


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    make_hash_cases()
    manifest = []
    for name, args in CASES:
        d = os.path.join(OUT, name)
        os.makedirs(d)
        a = [x.replace("OUT", os.path.join("wide", name)) for x in args]
        r = subprocess.run([REF_KAT] + a, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit(f"reference failed on {name}: {r.stdout[-800:]}")
        files = sorted(os.listdir(d))
        for f in files:
            p = os.path.join(d, f)
            if os.path.getsize(p) > 30_000:
                with open(p, "rb") as src, gzip.GzipFile(p + ".gz", "wb", mtime=0) as dst:
                    dst.write(src.read())
                os.unlink(p)
        manifest.append({"name": name, "args": args, "files": files})
        print(name, files)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()

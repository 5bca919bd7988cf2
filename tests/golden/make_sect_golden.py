#!/usr/bin/env python
"""Golden fixtures for `kat sect`: the files the UNMODIFIED reference binary
writes for a handful of commands over tests/golden/inputs, stored under tests/golden/sect/.
Run from the repo root where /root/reference exists:  make -C oracle ref && python tests/golden/make_sect_golden.py
Commands run with cwd = tests/golden and relative paths, as make_golden.py does."""
import gzip
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_KAT = os.path.join(ROOT, "oracle", "_ref", "kat")
OUT = os.path.join(HERE, "sect")

CASES = [
    ("sect_asm_vs_reads_k21", ["sect", "-m21", "-o", "OUT/s", "inputs/synth_c.fa", "inputs/synth_a.fq"]),
    ("sect_regions_gc_k17", ["sect", "-m17", "-E", "-F", "-g", "-M", "3", "-G", "40", "-o", "OUT/s", "inputs/synth_c.fa", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("sect_fastq_records_k13", ["sect", "-m13", "-o", "OUT/s", "inputs/synth_b.fq", "inputs/synth_a.fq"]),
    ("sect_edge_short_iupac_k5", ["sect", "-m5", "-g", "-E", "-o", "OUT/s", "inputs/edge1.fa", "inputs/edge1.fa"]),
    ("sect_multiline_fastq_nc_k5", ["sect", "-m5", "-N", "-o", "OUT/s", "inputs/edge2.fq", "inputs/edge2.fq"]),
    ("sect_no_count_stats_k27", ["sect", "-m27", "-n", "-o", "OUT/s", "inputs/long.fa", "inputs/synth_c.fa"]),
    # kat cold <assembly> <reads...>: contig-level read coverage and assembly copy number
    ("cold_asm_vs_reads_k21", ["cold", "-m21", "-o", "OUT/c", "inputs/synth_c.fa", "inputs/synth_a.fq", "inputs/synth_b.fq"]),
    ("cold_edge_k5", ["cold", "-m5", "-o", "OUT/c", "inputs/edge1.fa", "inputs/edge1.fa"]),
    ("cold_trim_k17", ["cold", "-m17", "--5ptrim", "4", "-H", "5000", "-o", "OUT/c", "inputs/synth_c.fa", "inputs/synth_a.fq"]),
]


def main():
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    manifest = []
    for name, args in CASES:
        d = os.path.join(OUT, name)
        os.makedirs(d)
        a = [x.replace("OUT", os.path.join("sect", name)) for x in args]
        r = subprocess.run([REF_KAT] + a, cwd=HERE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise SystemExit(f"reference failed on {name}: {r.stdout[-800:]}")
        files = sorted(os.listdir(d))
        for f in files:
            p = os.path.join(d, f)
            if os.path.getsize(p) > 100_000:
                with open(p, "rb") as src, gzip.GzipFile(p + ".gz", "wb", mtime=0) as dst:
                    dst.write(src.read())
                os.unlink(p)
        manifest.append({"name": name, "args": args, "files": files})
        print(name, files)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()

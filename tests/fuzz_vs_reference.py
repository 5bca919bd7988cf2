"""Differential fuzzing of the `kat` front-end against the UNMODIFIED reference binary (oracle/_ref/kat), beyond the
hist / gcp / comp option space of tests/test_host_frontend.py: sect with all its switches, cold, filter kmer, filter seq
(single and paired), gzip'd inputs with per-file 5' trims, and jellyfish hashes in both directions (reference-written
hash -> our LOAD mode, our `-d` dump -> the reference's LOAD mode).

One case = one random command run twice, each binary in its own copy of the same working directory (so every relative
path, and with it every header line, is the same): exit codes equal, the same set of output files, every file equal byte
for byte — jellyfish hashes by their records (`kat_jellyfish dump -c`, sorted), since their headers carry the command
line and a timestamp.  The reference needs no GPU; the front-end runs against whatever libkatgpu the environment gives it
(the real library on a GPU box, the test double of tests/hostcheck under LD_PRELOAD on a CPU-only machine).

As a module: `case(kind, seed, tmp_path)` asserts one case (used by tests/test_host_frontend.py with a few seeds).
As a script: `python tests/fuzz_vs_reference.py <kind|all> <first seed> <last seed>` runs a campaign on the CPU double
(KATFUZZ_GPUS=N: the front-end as a job of N devices; KATFUZZ_RAW=1: the double takes raw FASTQ blocks, so the
front-end's device-parsing route runs, in 4 KiB blocks; KATFUZZ_REAL=1 on a GPU box: no double, the real library)."""
import gzip
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = os.path.join(ROOT, "kat_b200", "bin", "kat")
REF = os.path.join(ROOT, "oracle", "_ref", "kat")
REF_JF = os.path.join(ROOT, "oracle", "_ref", "kat_jellyfish")
KINDS = ("sect", "cold", "fkmer", "fseq", "gz", "load", "group", "compgroups", "jfcount", "edgeopts", "edgeopts2")


def records(rng, n_rec, fastq, max_len=300, alphabet=b"ACGTacgtNRY-", p=(.22, .22, .22, .22, .02, .02, .02, .02, .01, .01, .01, .01)):
    """n_rec records, ragged lengths (a few long ones), single- or multi-line, FASTQ qualities with '@' / '+' / '>' in them."""
    out = bytearray()
    letters = np.frombuffer(alphabet, dtype=np.uint8)
    crlf = os.environ.get("KATFUZZ_CRLF") and rng.random() < 0.5  # (campaign option: Windows line ends in half of the files)
    for i in range(n_rec):
        n = int(rng.integers(0, 6000)) if rng.random() < 0.1 else int(rng.integers(0, max_len))
        seq = bytes(rng.choice(letters, n, p=p))
        width = int(rng.integers(1, 120)) if rng.random() < 0.4 else max(n, 1)
        lines = [seq[j:j + width] for j in range(0, n, width)] or [b""]
        head = [b"r%d/1 text", b"r%d", b"r%d\tx=1 y", b"r%d " + b"z" * 300][int(rng.integers(0, 4))] % i
        if fastq:
            q = bytes(rng.choice(np.frombuffer(b"I@+>#5", dtype=np.uint8), n))
            out += b"@" + head + b"\n" + b"\n".join(lines) + b"\n+\n" + (b"\n".join(q[j:j + width] for j in range(0, n, width)) or b"") + b"\n"
        else:
            out += b">" + head + b"\n" + b"\n".join(lines) + b"\n"
    return bytes(out).replace(b"\n", b"\r\n") if crlf else bytes(out)


def related(rng, src, n_rec, fastq, read_len=80):
    """Reads cut out of the bases of `src` (so that lookups find something), with a few substitutions."""
    bases = bytes(b for b in src if b in b"ACGT")
    out = bytearray()
    for i in range(n_rec):
        if len(bases) > read_len and rng.random() < 0.8:
            o = int(rng.integers(0, len(bases) - read_len))
            s = bytearray(bases[o:o + int(rng.integers(read_len // 4, read_len))])
            for _ in range(int(rng.integers(0, 3))):
                s[int(rng.integers(0, len(s)))] = b"ACGTN"[int(rng.integers(0, 5))]
            s = bytes(s)
        else:
            s = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(0, read_len))))
        out += (b"@q%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n") if fastq else (b">q%d\n" % i + s + b"\n")
    return bytes(out)


def _seq_files(rng, d, n, allow_gz=False):
    """n sequence files in d: a FASTA 'genome', then files related to it (mixed FASTA / FASTQ)."""
    names = []
    genome = records(rng, int(rng.integers(1, 8)), False, max_len=1500, alphabet=b"ACGTN", p=(.2475, .2475, .2475, .2475, .01))
    for i in range(n):
        fq = rng.random() < 0.5
        if i == 0:
            data, fq = genome, False
        elif rng.random() < 0.7:
            data = related(rng, genome, int(rng.integers(1, 60)), fq)
        else:
            data = records(rng, int(rng.integers(1, 12)), fq)
        name = "in%d.%s" % (i, "fq" if fq else "fa")
        if allow_gz and rng.random() < 0.5:
            name += ".gz"
            data = gzip.compress(data, mtime=0)
        open(os.path.join(d, name), "wb").write(data)
        names.append(name)
    return names


def _command(kind, rng, d):
    """-> (list of command lines to run in order, in directory d)."""
    k = int(rng.choice([3, 5, 7, 11, 17, 21, 33, 40]))
    flag = lambda f, p=0.3: [f] if rng.random() < p else []
    if kind == "sect":
        names = _seq_files(rng, d, int(rng.integers(2, 4)))
        seq = names[int(rng.integers(0, len(names)))]
        args = ["sect", "-m", str(k), "-H", "4096", "-o", "out/x"] + flag("-N") + flag("-n") + flag("-g", 0.5) + flag("-E", 0.5) + flag("-F", 0.5) + flag("-d", 0.2)
        if rng.random() < 0.5:
            args += ["-M", str(int(rng.integers(1, 5)))]
        if rng.random() < 0.5:
            args += ["-G", str(int(rng.integers(0, 40)))]
        if rng.random() < 0.3:
            args += ["--5ptrim", str(int(rng.integers(0, 6)))]
        if rng.random() < 0.3:
            args += ["-x", str(int(rng.integers(2, 50))), "-y", str(int(rng.integers(2, 50)))] + flag("-l", 0.5)
        return [args + [seq] + names[:int(rng.integers(1, len(names) + 1))]]
    if kind == "cold":
        names = _seq_files(rng, d, int(rng.integers(2, 4)))
        args = ["cold", "-m", str(k), "-H", "4096", "-o", "out/c"]
        if rng.random() < 0.4:
            args += ["--5ptrim", str(int(rng.integers(0, 6)))]
        if rng.random() < 0.3:
            args += ["-x", str(int(rng.integers(2, 50))), "-y", str(int(rng.integers(2, 50)))]
        return [args + [names[0]] + names[1:]]
    if kind == "fkmer":
        names = _seq_files(rng, d, int(rng.integers(1, 3)))
        lo_c, lo_g = int(rng.integers(0, 4)), int(rng.integers(0, k))
        args = ["filter", "kmer", "-m", str(k), "-H", "4096", "-o", "out/k", "-c", str(lo_c), "-d", str(lo_c + int(rng.integers(0, 6))),
                "-g", str(lo_g), "-h", str(lo_g + int(rng.integers(0, k)))] + flag("-i") + flag("-s", 0.5) + flag("-N")
        return [args + names]
    if kind == "fseq":
        names = _seq_files(rng, d, 2)
        fq = rng.random() < 0.5
        n_rec = int(rng.integers(1, 40))
        ext = "fq" if fq else "fa"
        src = open(os.path.join(d, names[0]), "rb").read()
        open(os.path.join(d, "s1." + ext), "wb").write(related(rng, src, n_rec, fq) if rng.random() < 0.6 else records(rng, n_rec, fq))
        args = ["filter", "seq", "-m", str(k), "-H", "4096", "-o", "out/f", "--stats", "-T", str(float(rng.choice([0.0, 0.1, 0.3, 0.5, 0.9, 1.0])))]
        args += flag("-i") + flag("-s", 0.6) + flag("-N") + ["--seq", "s1." + ext]
        if rng.random() < 0.5:  # paired: the second file has the same number of records — or, sometimes, fewer / more
            n2 = n_rec if rng.random() < 0.8 else max(1, n_rec + int(rng.integers(-2, 3)))
            open(os.path.join(d, "s2." + ext), "wb").write(related(rng, src, n2, fq) if rng.random() < 0.6 else records(rng, n2, fq))
            args += ["--seq2", "s2." + ext]
        return [args + names[:int(rng.integers(1, 3))]]
    if kind == "gz":
        names = _seq_files(rng, d, 3, allow_gz=True)
        tool = ["hist", "gcp", "comp"][int(rng.integers(0, 3))]
        trims = lambda n: ",".join(str(int(rng.integers(0, 5))) for _ in range(n))
        threads = ["-t", str(int(rng.integers(1, 4)))]
        if tool == "comp":
            args = ["comp", "-m", str(k), "-H", "4096", "-I", "4096", "-o", "out/c"] + threads + flag("-N") + flag("-O") + flag("-h") + flag("-n")
            if rng.random() < 0.5:
                args += ["--d1_5ptrim", trims(1), "--d2_5ptrim", trims(1)]
            return [args + names[:2]]
        n_in = int(rng.integers(1, 4))
        args = [tool, "-m", str(k), "-H", "4096", "-o", "out/o"] + threads + flag("-N")
        if rng.random() < 0.6:
            args += ["--5ptrim", trims(n_in if rng.random() < 0.7 else 1)]
        return [args + names[:n_in]]
    if kind == "group":
        # groups of 2-5 files read by 1-4 parser streams (-t): regular files, files whose last FASTQ record is broken
        # ("Invalid fastq sequence" ends the stream that read it, the other streams go on), files in no supported format
        # (fatal among the files opened while the parser is built, otherwise the end of one stream), empty files; with a
        # single stream also files that end inside a record (the seq_len they leave behind reaches the next file; with
        # several live streams the reference races for which one that is)
        threads = int(rng.integers(1, 5))
        n_files = int(rng.integers(2, 6))
        one_stream = min(threads, n_files) == 1
        names = []
        for i in range(n_files):
            fq = rng.random() < 0.6
            data = records(rng, int(rng.integers(1, 8)), fq, max_len=120, alphabet=b"ACGTN", p=(.24, .24, .24, .24, .04))
            if rng.random() < 0.15:  # more than one 4 KiB buffer
                data += records(rng, int(rng.integers(20, 60)), fq, max_len=400, alphabet=b"ACGTN", p=(.24, .24, .24, .24, .04))
            if fq and not one_stream and data.endswith(b"\n\n+\n\n"):  # an empty last record leaves seq_len behind: single-stream groups only
                data += b"@z\nACGT\n+\nIIII\n"
            x = rng.random()
            if x < 0.08:
                data = b""
            elif x < 0.18 and (one_stream or i < threads):  # (which of several live streams opens a later file is a race)
                data = b"#" + data
            elif fq and x < 0.45:
                data += [b"@t\nACGTACGTCC\n+\nIIII", b"@t\nACGTTGCATG\n+\nIIIIIII\n", b"@t\nGGGTTTCCCA\n+\nIIIIIIIIIIIII\n"][int(rng.integers(0, 3))]
            elif fq and x < 0.6 and one_stream:
                data += [b"@t\n\n+\n\n", b"@t\nACGTACGT\n", b"@t\nACGTAC"][int(rng.integers(0, 3))]
            names.append("g%d.%s" % (i, "fq" if fq else "fa"))
            open(os.path.join(d, names[-1]), "wb").write(data)
        tool = ["hist", "gcp"][int(rng.integers(0, 2))]
        return [[tool, "-m", str(int(rng.choice([3, 5, 9]))), "-H", "4096", "-t", str(threads), "-o", "out/o"] + flag("-N") + names]
    if kind == "compgroups":
        # comp whose inputs are GROUPS of files: space-separated names and glob patterns in one argument, a trim per file of
        # the group, 2 or 3 inputs, -t
        names = _seq_files(rng, d, 5, allow_gz=True)
        rng.shuffle(names)
        n_in = 3 if rng.random() < 0.3 else 2
        cuts = sorted(rng.choice(np.arange(1, 5), n_in - 1, replace=False).tolist())
        groups = [names[a:b] for a, b in zip([0] + cuts, cuts + [5])]
        args = ["comp", "-m", str(k), "-H", "4096", "-I", "4096", "-J", "4096", "-o", "out/c", "-t", str(int(rng.integers(1, 4)))]
        args += flag("-N") + flag("-O") + flag("-h") + flag("-n")
        if n_in == 3:
            args += flag("-P")
        if rng.random() < 0.5:
            args += ["-i", str(int(rng.integers(2, 40))), "-j", str(int(rng.integers(2, 40)))]
        if rng.random() < 0.4:
            args += ["-x", str(float(rng.choice([0.5, 2.0]))), "-y", str(float(rng.choice([0.25, 1.5])))]
        for opt, g in zip(("--d1_5ptrim", "--d2_5ptrim"), groups):
            if rng.random() < 0.4:
                args += [opt, ",".join(str(int(rng.integers(0, 5))) for _ in range(len(g) if rng.random() < 0.7 else 1))]
        as_arg = lambda g: " ".join(g) if rng.random() < 0.7 or len(g) == 1 else "in[%s].*" % "".join(n[2] for n in g)
        return [args + [as_arg(g) for g in groups]]
    if kind == "jfcount":
        # hashes written by the unmodified `jellyfish count` with its own options (table size, counter bits, counter bytes in
        # the file, -L / -U bounds, canonical or not) as inputs of hist / gcp / comp / sect / filter kmer (LOAD mode)
        # (k >= 11 keeps the counts small: the reference's LOAD mode ignores add()'s result and loses records when the chained
        # large-count entries of a narrow counter field fill its 2x-sized table, see "load" below; a hash without records makes
        # it fail with "Invalid matrix size 0x<key bits>", exit code 5, which the front-end reproduces; tables started very small,
        # `-s 64`, put a short reprobe limit into the header, with which the same silent loss happens at any load)
        names = _seq_files(rng, d, 3)
        k = max(k, 11)
        jf = ["!jellyfish", "count", "-m", str(k), "-s", str(int(rng.choice([1000, 20000, 300000]))), "-t", str(int(rng.integers(1, 3))), "-o", "x.jf"]
        jf += flag("-C", 0.6)
        if rng.random() < 0.5:
            jf += ["-c", str(int(rng.integers(2, 9)))]
        if rng.random() < 0.5:
            jf += ["--out-counter-len", str(int(rng.integers(1, 5)))]
        if rng.random() < 0.3:
            jf += ["-L", str(int(rng.integers(1, 4)))]
        if rng.random() < 0.3:
            jf += ["-U", str(int(rng.integers(3, 40)))]
        jf += names[:int(rng.integers(1, 3))]
        second = [["hist", "-o", "out/h", "-l", str(int(rng.integers(1, 4))), "-h", str(int(rng.integers(4, 300))), "x.jf"],
                  ["gcp", "-o", "out/g", "-y", str(int(rng.integers(2, 300))), "x.jf"],
                  ["comp", "-m", str(k), "-H", "4096", "-o", "out/c", "-i", "50", "-j", "50"] + flag("-N") + flag("-h") + ["x.jf", names[2]],
                  ["comp", "-m", str(k), "-H", "4096", "-o", "out/c", "-i", "50", "-j", "50"] + flag("-O") + [names[2], "x.jf"],
                  ["sect", "-o", "out/s", "-m", str(k), names[0], "x.jf"],
                  ["filter", "kmer", "-o", "out/k", "-c", str(int(rng.integers(0, 3))), "-d", str(int(rng.integers(3, 300))), "-g", "0", "-h", str(k), "-s", "x.jf"]]
        return [jf, second[int(rng.integers(0, len(second)))]]
    if kind == "edgeopts":
        # option values at and beyond their edges: bounds of 0 and 1, more bins than counts, single-cell matrices, scales
        # far from 1, the shortest and the longest mers, a hash size of 1 ("Invalid matrix size", exit code 5).  (Hash sizes of
        # a few slots are left out: with 3-mers in 128 slots the reference itself loses a carry of 128 while it grows the table.)
        names = _seq_files(rng, d, 3)
        k = int(rng.choice([1, 2, 3, 4, 16, 31, 32, 33, 63, 64]))
        pick = lambda xs: str(xs[int(rng.integers(0, len(xs)))])
        tool = ["hist", "gcp", "comp"][int(rng.integers(0, 3))]
        if tool == "hist":
            args = ["hist", "-m", str(k), "-H", pick([1, 4096]), "-o", "out/h", "-l", pick([0, 1, 2, 3, 50]), "-h", pick([0, 1, 2, 3, 49, 100000]),
                    "-i", pick([1, 2, 3, 1000])] + flag("-N") + names[:int(rng.integers(1, 3))]
        elif tool == "gcp":
            args = ["gcp", "-m", str(k), "-H", pick([1, 4096]), "-o", "out/g", "-x", pick([0.001, 0.01, 0.3333, 1, 3.7, 1000]), "-y", pick([1, 2, 3, 1000, 65535])]
            args += flag("-N") + names[:int(rng.integers(1, 3))]
        else:
            args = ["comp", "-m", str(k), "-H", pick([1, 4096]), "-I", pick([1, 4096]), "-o", "out/c", "-i", pick([1, 2, 3, 1001]), "-j", pick([1, 2, 3, 500]),
                    "-x", pick([0.001, 0.3333, 1, 2.5, 100]), "-y", pick([0.01, 1, 7.5])] + flag("-N") + flag("-O") + flag("-h") + flag("-n") + names[:2]
        return [args]
    if kind == "edgeopts2":
        # the same for sect / cold / filter: repeat bounds of 0 and in the wrong order, one-bin matrices, thresholds outside
        # [0, 1], count and GC bounds of 0, beyond every count and in the wrong order, the shortest and longest mers
        names = _seq_files(rng, d, 3)
        k = int(rng.choice([1, 2, 3, 16, 31, 32, 33, 64]))
        pick = lambda xs: str(xs[int(rng.integers(0, len(xs)))])
        tool = ["sect", "cold", "fseq", "fkmer"][int(rng.integers(0, 4))]
        if tool == "sect":
            args = ["sect", "-m", str(k), "-H", "4096", "-o", "out/s", "-M", pick([0, 1, 2, 1000000]), "-G", pick([0, 1, 2, 5000000000]),
                    "-x", pick([0, 1, 2, 1001]), "-y", pick([0, 1, 2, 1001])] + flag("-l") + flag("-g", 0.5) + flag("-E", 0.6) + flag("-F", 0.6) + flag("-n") + flag("-N")
            return [args + [names[int(rng.integers(0, 3))]] + names[:int(rng.integers(1, 3))]]
        if tool == "cold":
            return [["cold", "-m", str(k), "-H", pick([1, 4096]), "-o", "out/c", "-x", pick([0, 1, 1001]), "-y", pick([0, 1, 1001])] + [names[0]] + names[1:]]
        if tool == "fseq":
            args = ["filter", "seq", "-m", str(k), "-H", "4096", "-o", "out/f", "--stats", "-T", pick([-0.5, 0, 1e-9, 0.999999, 1, 1.5])] + flag("-i") + flag("-s", 0.6) + flag("-N")
            return [args + ["--seq", names[1]] + names[:1]]
        args = ["filter", "kmer", "-m", str(k), "-H", "4096", "-o", "out/k", "-c", pick([0, 1, 5, 18446744073709551615]), "-d", pick([0, 1, 4, 100000, 18446744073709551615]),
                "-g", pick([0, 1, 30, 65535]), "-h", pick([0, 1, 29, 64, 65535])] + flag("-i") + flag("-s", 0.5) + flag("-N")
        return [args + names[:int(rng.integers(1, 3))]]
    if kind == "load":
        # (k >= 5: the reference's LOAD mode sizes its table at 2x the records and ignores add()'s result, lib/src/jellyfish_helper.cc:
        # 160-172 — with the 32 canonical 3-mers and counts beyond the 7-bit field, which take a second slot each, it
        # silently loses a record; a defect of tiny tables that the front-end does not copy)
        k = max(k, 5)
        names = _seq_files(rng, d, 2)
        canon = flag("-N")
        first = ["hist", "-m", str(k), "-H", "4096", "-d", "-o", "out/h"] + canon + names[:int(rng.integers(1, 3))]
        jf = "out/h-hash.jf%d" % k
        second = [["hist", "-o", "out/h2", "-l", str(int(rng.integers(1, 4))), "-h", str(int(rng.integers(4, 30))), jf],
                  ["gcp", "-o", "out/g2", "-y", str(int(rng.integers(2, 40))), jf],
                  ["comp", "-m", str(k), "-H", "4096", "-o", "out/c2", "-i", "30", "-j", "30"] + canon + [jf, names[1]],
                  ["sect", "-o", "out/s2", "-m", str(k), names[0], jf]][int(rng.integers(0, 4))]
        return [first, second]
    raise ValueError(kind)


class ReferenceHung(RuntimeError):
    """The reference binary did not end (it can spin in its own table growth on degenerate inputs): not a case."""


def _run(exe, cmds, cwd, env):
    os.makedirs(os.path.join(cwd, "out"), exist_ok=True)
    codes = []
    for args in cmds:
        if args[0] == "!jellyfish":  # a preparation step, run with the unmodified jellyfish on both sides
            r = subprocess.run([REF_JF] + args[1:], cwd=cwd, capture_output=True, text=True, env={k: v for k, v in env.items() if k != "LD_PRELOAD"}, timeout=300)
            codes.append(r.returncode)
            continue
        for attempt in range(6):
            try:
                r = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, env=env, timeout=150)
            except subprocess.TimeoutExpired:
                if exe == REF:
                    raise ReferenceHung(" ".join(args))
                raise
            # the reference reads an uninitialised terminal width when stdout is no tty and may abort on it; with -t > 1 its
            # worker threads push_back into one vector unlocked (src/histogram.cc:198) and it may crash (backtrace, exit code 1)
            if r.returncode in (0, 4, 5) or exe != REF:
                break
        codes.append(r.returncode)
    files = {}
    for f in sorted(os.listdir(os.path.join(cwd, "out"))):
        p = os.path.join(cwd, "out", f)
        if ".jf" in f:
            j = subprocess.run([REF_JF, "dump", "-c", p], capture_output=True, text=True, env=env)
            files[f] = ("jellyfish records", j.returncode, sorted(j.stdout.splitlines()))
        else:
            files[f] = open(p, "rb").read()
    return codes, files


def on_a_job(cmds):
    """What `kat --gpus N` (N > 1) takes: no jellyfish hashes in or out, one-word k-mers, comp inputs all canonical."""
    for c in cmds:
        k = int(c[c.index("-m") + 1]) if "-m" in c else 27
        if c[0] == "!jellyfish" or "-d" in c[:-1] and c[0] != "filter" or any(a.endswith(".jf") for a in c) or k > 32:
            return False
        if c[0] == "comp" and {"-N", "-O", "-P"} & set(c):
            return False
        if c[:2] == ["filter", "kmer"]:
            pass
    return True


def case(kind, seed, tmp_path, our_env=None, gpus=1):
    """gpus > 1: the front-end side runs as a job of that many devices (KATGPU_GPUS; on a CPU-only machine the devices of the
    test double, KATGPU_DOUBLE_DEVICES); commands a job does not take are skipped (returns False)."""
    rng = np.random.default_rng([KINDS.index(kind), seed])
    tmp_path = str(tmp_path)
    dirs = {who: os.path.join(tmp_path, who) for who in ("ref", "our")}
    os.makedirs(dirs["ref"])
    cmds = _command(kind, rng, dirs["ref"])
    if gpus > 1:
        if not on_a_job(cmds):
            return False
        our_env = dict(os.environ if our_env is None else our_env, KATGPU_GPUS=str(gpus), KATGPU_DOUBLE_DEVICES=str(gpus))
    shutil.copytree(dirs["ref"], dirs["our"])
    no_preload = {k: v for k, v in os.environ.items() if k != "LD_PRELOAD"}
    ref = _run(REF, cmds, dirs["ref"], no_preload)
    our = _run(KAT, cmds, dirs["our"], dict(os.environ if our_env is None else our_env))
    threaded = any(a == "-t" and c[i + 1] != "1" for c in cmds for i, a in enumerate(c[:-1]))
    for again in range(3):  # the same unlocked push_back can also lose a thread's share without crashing: such a run is not the reference's answer
        if not threaded or (our[0] == ref[0] and our[1] == ref[1]):
            break
        shutil.rmtree(os.path.join(dirs["ref"], "out"))
        ref = _run(REF, cmds, dirs["ref"], no_preload)
    what = "%s seed %d: %s" % (kind, seed, " && ".join(" ".join(c) for c in cmds))
    assert our[0] == ref[0], what + ": exit codes %r, reference %r" % (our[0], ref[0])
    assert sorted(our[1]) == sorted(ref[1]), what + ": files %r, reference %r" % (sorted(our[1]), sorted(ref[1]))
    for f in ref[1]:
        assert our[1][f] == ref[1][f], what + ": " + f + " differs"
    if kind == "load" and ref[0][0] == 0:  # and the other direction: the reference loads the hash WE wrote
        jf = cmds[1][-1] if cmds[1][0] != "comp" else cmds[1][-2]
        for who in dirs:
            os.makedirs(os.path.join(dirs[who], "back"))
        shutil.copy(os.path.join(dirs["our"], jf), os.path.join(dirs["ref"], "ours.jf"))
        shutil.copy(os.path.join(dirs["ref"], jf), os.path.join(dirs["our"], "ours.jf"))  # same name, the reference's own file
        back = []
        for who in ("ref", "our"):
            for attempt in range(4):
                r = subprocess.run([REF, "hist", "-o", "back/h", "ours.jf"], cwd=dirs[who], capture_output=True, text=True, env=no_preload)
                if r.returncode >= 0:
                    break
            back.append((r.returncode, open(os.path.join(dirs[who], "back", "h"), "rb").read() if r.returncode == 0 else b""))
        assert back[0] == back[1], what + ": the reference reads our dump differently from its own"


if __name__ == "__main__":
    kinds = KINDS if sys.argv[1] == "all" else (sys.argv[1],)
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    if os.environ.get("KATFUZZ_REAL"):  # on a GPU box: the real libkatgpu answers
        env = dict(os.environ)
    else:
        double = os.path.join(ROOT, "tests", "hostcheck", "_build", "libkatgpu_double.so")
        assert subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostcheck")], capture_output=True).returncode == 0
        env = dict(os.environ, LD_PRELOAD=double, KAT_GPU_PARSE="1")
    gpus = int(os.environ.get("KATFUZZ_GPUS", "1"))
    if os.environ.get("KATFUZZ_RAW"):  # the double takes raw FASTQ blocks like the real library: the front-end's raw route, in 4 KiB blocks
        env.update(KATGPU_DOUBLE_FASTQ="1", KAT_RAW_BLOCK="4096")
    bad = []
    for kind in kinds:
        for seed in range(lo, hi):
            d = "/tmp/katfuzz_%s_%d_%d" % (kind, seed, os.getpid())
            shutil.rmtree(d, ignore_errors=True)
            os.makedirs(d)
            try:
                case(kind, seed, d, env, gpus)
                shutil.rmtree(d, ignore_errors=True)
            except AssertionError as e:
                bad.append((kind, seed))
                print("FAIL", str(e)[:600], "(kept " + d + ")", flush=True)
            except Exception as e:  # noqa: BLE001 — a campaign reports and goes on
                bad.append((kind, seed))
                print("ERROR", kind, seed, repr(e)[:400], "(kept " + d + ")", flush=True)
    print("done:", kinds, lo, hi, "failures:", bad, flush=True)

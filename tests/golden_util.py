"""Helpers to read the golden fixtures (tests/golden, produced by make_golden.py)."""
import gzip
import json
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)


def wide_cases():
    """hist / gcp / comp commands at k = 33..64 (tests/golden/wide, make_wide_golden.py)."""
    with open(os.path.join(GOLDEN, "wide", "cases.json")) as f:
        return [dict(c, dir="wide") for c in json.load(f) if c["args"][0] in ("hist", "gcp", "comp")]


def expected_bytes(case, fname, sub="expected"):
    p = os.path.join(GOLDEN, sub, case, fname)
    if os.path.exists(p):
        return open(p, "rb").read()
    return gzip.open(p + ".gz", "rb").read()


def parse_cli(args):
    """Split a golden command into (mode, options dict, positional list)."""
    mode, rest = args[0], args[1:]
    takes = {"-o", "-t", "-l", "-h", "-i", "-j", "-x", "-y", "-m", "-H", "-I", "-J", "--5ptrim", "--d1_5ptrim", "--d2_5ptrim"}
    if mode == "comp":
        takes = takes - {"-h"}
    opts, pos, i = {}, [], 0
    while i < len(rest):
        a = rest[i]
        if a.startswith("-m") and len(a) > 2:
            opts["-m"] = a[2:]
        elif a.startswith("-t") and len(a) > 2:
            opts["-t"] = a[2:]
        elif a in takes:
            opts[a] = rest[i + 1]
            i += 1
        elif a.startswith("-"):
            opts[a] = True
        else:
            pos.append(a)
        i += 1
    return mode, opts, pos


def parse_hist_text(b):
    lab, val = [], []
    for line in b.decode().splitlines():
        if line.startswith("#") or not line.strip():
            continue
        x, y = line.split()
        lab.append(int(x)); val.append(int(y))
    return np.array(lab, dtype=np.uint64), np.array(val, dtype=np.uint64)


def parse_mx_text(b):
    hdr, rows = {}, []
    for line in b.decode().splitlines():
        if line.startswith("#"):
            if ":" in line:
                k, v = line[1:].split(":", 1)
                hdr[k.strip()] = v.strip()
            continue
        if line.strip():
            rows.append(np.array(line.split(), dtype=np.uint64))
    return hdr, np.vstack(rows)


def parse_stats_text(b):
    """.stats -> (13 integer counters in katgpu_comp_counters order (hash3 = 0 if absent), 10 distance strings)."""
    t = b.decode()
    def grab(section, label):
        m = re.search(re.escape(section) + r".*?" + re.escape(label) + r"\s*(\d+)", t, re.S)
        return int(m.group(1)) if m else 0
    c = {
        "hash1_total": grab("Total K-mers in:", "- Hash 1:"), "hash2_total": grab("Total K-mers in:", "- Hash 2:"),
        "hash1_distinct": grab("Distinct K-mers in:", "- Hash 1:"), "hash2_distinct": grab("Distinct K-mers in:", "- Hash 2:"),
        "hash1_only_total": grab("Total K-mers only found in:", "- Hash 1:"),
        "hash2_only_total": grab("Total K-mers only found in:", "- Hash 2:"),
        "hash1_only_distinct": grab("Distinct K-mers only found in:", "- Hash 1:"),
        "hash2_only_distinct": grab("Distinct K-mers only found in:", "- Hash 2:"),
        "shared_hash1_total": grab("Shared K-mers:", "Total shared found in hash 1:"),
        "shared_hash2_total": grab("Shared K-mers:", "Total shared found in hash 2:"),
        "shared_distinct": grab("Shared K-mers:", "Distinct shared K-mers:"),
    }
    m3 = re.search(r"Total K-mers in:.*?- Hash 3: (\d+)", t, re.S)
    c["hash3_total"] = int(m3.group(1)) if m3 else 0
    m3 = re.search(r"Distinct K-mers in:.*?- Hash 3: (\d+)", t, re.S)
    c["hash3_distinct"] = int(m3.group(1)) if m3 else 0
    dist = re.findall(r"distance: (\S+)", t)
    return c, dist

"""Small deterministic sequence generators for the parity tests (numpy only)."""
import numpy as np

ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def random_stream(n, seed, n_rate=0.01, lower_rate=0.05, junk_rate=0.002):
    """Base stream with N separators, lower-case bases and a few non-base bytes
    (IUPAC codes, '-', '\\r', digits) — everything mer_dna::code treats as a break."""
    rng = np.random.default_rng(seed)
    s = ALPHA[rng.integers(0, 4, n)].copy()
    low = rng.random(n) < lower_rate
    s[low] |= 0x20
    s[rng.random(n) < n_rate] = ord("N")
    junk = np.frombuffer(b"RYKMSWn-\r.0*", dtype=np.uint8)
    jm = rng.random(n) < junk_rate
    s[jm] = junk[rng.integers(0, len(junk), int(jm.sum()))]
    return s


def genome_reads(genome_len, n_reads, read_len, seed, sub_rate=0.01):
    """Reads sampled from one random genome (both strands) -> high-multiplicity k-mers."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, genome_len).astype(np.uint8)
    pos = rng.integers(0, genome_len - read_len + 1, n_reads)
    idx = pos[:, None] + np.arange(read_len)[None, :]
    r = g[idx]
    rev = rng.random(n_reads) < 0.5
    r[rev] = 3 - r[rev][:, ::-1]
    sub = rng.random(r.shape) < sub_rate
    r[sub] = (r[sub] + rng.integers(1, 4, int(sub.sum()))) & 3
    out = np.full((n_reads, read_len + 1), ord("N"), dtype=np.uint8)
    out[:, :read_len] = ALPHA[r]
    return out.reshape(-1)


def low_complexity(n_reads, read_len, seed):
    """poly-A / dinucleotide repeats: a handful of k-mers with huge counts."""
    rng = np.random.default_rng(seed)
    out = np.full((n_reads, read_len + 1), ord("N"), dtype=np.uint8)
    for i in range(n_reads):
        m = rng.integers(0, 3)
        if m == 0:
            out[i, :read_len] = ord("A")
        elif m == 1:
            out[i, :read_len] = np.resize(np.frombuffer(b"AC", dtype=np.uint8), read_len)
        else:
            out[i, :read_len] = np.resize(np.frombuffer(b"ACGGT", dtype=np.uint8), read_len)
    return out.reshape(-1)


def encode_kmer(s):
    m = 0
    for ch in s.upper():
        m = (m << 2) | "ACGT".index(ch)
    return m


SOUP_TOKENS = [b"@", b"+", b">", b"\n", b"\n", b"\n", b"\r\n", b"ACGT", b"A", b"C", b"N", b"acgt", b"IIII", b"#", b" ", b"\n\n", b"GGGGGG",
               b"TTTTTTT", b"@r1\n", b"+\n", b">c1\n", b"ACGTACGTAC\n", b"IIIIIIIIII\n"]


def token_soup(rng, always_marker=True):
    """Adversarial sequence-file images: record markers, line ends, bases, quality-like bytes and blanks in random order
    behind a FASTA or FASTQ first line — mostly malformed records (what the parsers do with them is part of the contract)."""
    first = [b"@r0\n", b">c0\n"][int(rng.integers(0, 2))] if always_marker or rng.random() < 0.9 else b""
    return first + b"".join(SOUP_TOKENS[int(rng.integers(0, len(SOUP_TOKENS)))] for _ in range(int(rng.integers(1, 40))))

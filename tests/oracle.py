"""ctypes front for oracle/libkatoracle.so — the CPU checker.

Test infrastructure only (see the header of oracle/kat_oracle.c).  Imported by
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by the
product package `kat_b200`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_KAT = os.path.join(REF_DIR, "kat")
REF_JELLYFISH = os.path.join(REF_DIR, "kat_jellyfish")
REF_GENSEQ = os.path.join(REF_DIR, "generate_sequence")

CC_NAMES = [
    "hash1_total", "hash2_total", "hash3_total",
    "hash1_distinct", "hash2_distinct", "hash3_distinct",
    "hash1_only_total", "hash2_only_total",
    "hash1_only_distinct", "hash2_only_distinct",
    "shared_hash1_total", "shared_hash2_total", "shared_distinct",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = os.path.join(ORACLE_DIR, "libkatoracle.so")
    src = os.path.join(ORACLE_DIR, "kat_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libkatoracle.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    u8p, u64p, vp = C.c_void_p, C.c_void_p, C.c_void_p
    L.oracle_parse_file.restype = C.c_long
    L.oracle_parse_file.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, u8p]
    L.oracle_parse_file_state.restype = C.c_long
    L.oracle_parse_file_state.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, u8p, C.POINTER(C.c_size_t)]
    L.oracle_parse_file_state2.restype = C.c_long
    L.oracle_parse_file_state2.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, u8p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    L.oracle_table_new.restype = vp
    L.oracle_table_new.argtypes = [C.c_int, C.c_int]
    L.oracle_table_free.argtypes = [vp]
    L.oracle_table_add.argtypes = [vp, u8p, C.c_size_t]
    L.oracle_table_finish.argtypes = [vp]
    L.oracle_table_distinct.restype = C.c_uint64
    L.oracle_table_distinct.argtypes = [vp]
    L.oracle_table_total.restype = C.c_uint64
    L.oracle_table_total.argtypes = [vp]
    L.oracle_table_export.argtypes = [vp, u64p, u64p]
    L.oracle_lookup.restype = C.c_uint64
    L.oracle_lookup.argtypes = [vp, C.c_uint64, C.c_int]
    L.oracle_table_export_wide.argtypes = [vp, u64p, u64p, u64p]
    L.oracle_lookup_wide.restype = C.c_uint64
    L.oracle_lookup_wide.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
    L.oracle_table_add_pairs.argtypes = [vp, u64p, u64p, u64p, C.c_size_t]
    L.oracle_hist_geometry.argtypes = [C.c_uint64, C.c_uint64, u64p, u64p, u64p]
    L.oracle_hist.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, u64p]
    L.oracle_gcp.argtypes = [vp, C.c_double, C.c_uint32, u64p]
    L.oracle_comp.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.c_uint32, C.c_uint32,
                              u64p, u64p, u64p, u64p, u64p, u64p]
    L.oracle_distances.argtypes = [u64p, u64p, C.c_size_t, C.c_void_p]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def parse_file(data: bytes, k: int, trim5p: int = 0, carry=None, last=None):
    """a1: file image -> (base stream bytes, status). status 0 ok, -1 unsupported
    format, -2 invalid fastq (stream truncated as the reference does).  `carry`: a one-element list holding the
    parser's seq_len state, which the reference keeps from one file of a group to the next (see kat_oracle.c).
    `last`: a two-element list that receives where the last 4 KiB buffer's bytes start in the stream and the seam before it."""
    src = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(len(data) + 2, dtype=np.uint8)
    st = C.c_size_t(carry[0] if carry else 0)
    lb = (C.c_size_t * 2)(0, 0)
    r = lib().oracle_parse_file_state2(_ptr(src) if len(data) else None, len(data), k, trim5p, _ptr(out), C.byref(st), lb)
    if carry is not None:
        carry[0] = st.value
    if last is not None:
        last[0], last[1] = lb[0], lb[1]
    if r == -1:
        return b"", -1
    if r <= -2:
        return out[: -(r + 2)].tobytes(), -2
    return out[:r].tobytes(), 0


class UnsupportedFormat(RuntimeError):
    """A file opened while the reference builds its parser starts with neither '>' nor '@': the tool fails, exit code 5."""


def parse_files(files, k, trim5p=None, threads=1):
    """Whole input group -> list of base-stream chunks to count (no k-mer spans two chunks).

    The reference reads a group through S = min(files, threads) parser streams (lib/src/jellyfish_helper.cc:231-233).
    Files are handed out in order: the first S while the parser is built (an empty file is skipped by the stream that
    opened it, which takes the next one too; "Unsupported format" there is fatal), later ones to the stream that reaches
    the end of its file first (modelled by bytes read).  A stream keeps its seq_len from file to file, and an exception
    inside produce() — "Invalid fastq sequence", "Unsupported format" on a later file — ends THAT stream only
    (JFI/cooperative_pool2.hpp:249-260); with one stream, the rest of the group is lost.  The second kind also costs the
    stream the last 4 KiB buffer of the file it had just finished (kat_oracle.c, oracle_parse_file_state2)."""
    trim5p = trim5p or [0] * len(files)
    if len(trim5p) == 1 and len(files) > 1:
        trim5p = trim5p * len(files)
    chunks = []
    n_streams = max(1, min(len(files), max(1, threads)))
    alive, carry, clock, has_file = [True] * n_streams, [[0] for _ in range(n_streams)], [0] * n_streams, [False] * n_streams
    held = [None] * n_streams  # (stream of the file a stream has just finished, where its last 4 KiB buffer starts): counted whole once
    built = 0                  # the stream has opened its next file (or none is left), without that buffer when that file is unsupported
    for path, t in zip(files, trim5p):
        if isinstance(path, (bytes, bytearray)):
            data = bytes(path)
            size = len(data)
        else:
            import gzip
            with open(path, "rb") as f:
                data = f.read()
            size = len(data)
            if data[:2] == b"\x1f\x8b":
                data = gzip.decompress(data)
        ctor = built < n_streams
        if ctor:
            s = built
        else:
            live = [j for j in range(n_streams) if alive[j]]
            if not live:
                break
            s = min(live, key=lambda j: (clock[j], j))
        clock[s] += size
        if not data:
            continue
        if ctor and not has_file[s]:
            has_file[s] = True
            built += 1
        last = [0, 0]
        seq, st = parse_file(data, k, t, carry[s], last)
        if st == -1:
            if ctor:
                raise UnsupportedFormat("Unsupported format")
            alive[s] = False
            if held[s] is not None:
                chunks.append(held[s][0][:held[s][1]])
                held[s] = None
            continue
        if held[s] is not None:
            chunks.append(held[s][0])
            held[s] = None
        if st == -2:
            alive[s] = False
            chunks.append(seq)
            continue
        held[s] = (seq, last[0])
    chunks += [h[0] for h in held if h is not None]
    return [c for c in chunks if c]


class Table:
    """Exact k-mer multiset (sorted keys + counts)."""

    def __init__(self, k, canonical=True):
        self.k, self.canonical = k, bool(canonical)
        self.h = lib().oracle_table_new(k, int(canonical))
        if not self.h:
            raise ValueError("k out of range")
        self._done = False

    def add(self, bases: bytes):
        a = np.frombuffer(bases, dtype=np.uint8)
        if len(a):
            lib().oracle_table_add(self.h, _ptr(a), len(a))
        return self

    def finish(self):
        if not self._done:
            lib().oracle_table_finish(self.h)
            self._done = True
        return self

    @classmethod
    def from_chunks(cls, chunks, k, canonical=True):
        t = cls(k, canonical)
        for c in chunks:
            t.add(c)
        return t.finish()

    @classmethod
    def from_files(cls, files, k, canonical=True, trim5p=None, threads=1):
        return cls.from_chunks(parse_files(files, k, trim5p, threads), k, canonical)

    @property
    def distinct(self):
        return int(lib().oracle_table_distinct(self.h))

    @property
    def total(self):
        return int(lib().oracle_table_total(self.h))

    def export(self):
        """k <= 32: (keys, counts); k > 32: (keys_hi, keys_lo, counts) — sorted by key."""
        n = self.distinct
        keys = np.empty(n, dtype=np.uint64)
        counts = np.empty(n, dtype=np.uint64)
        if self.k > 32:
            hi = np.empty(n, dtype=np.uint64)
            lib().oracle_table_export_wide(self.h, _ptr(hi), _ptr(keys), _ptr(counts))
            return hi, keys, counts
        lib().oracle_table_export(self.h, _ptr(keys), _ptr(counts))
        return keys, counts

    def add_pairs(self, keys, counts, keys_hi=None):
        """count[key] += count records (HashLoader::loadHash); call before finish()."""
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        counts = np.ascontiguousarray(counts, dtype=np.uint64)
        hi = None if keys_hi is None else np.ascontiguousarray(keys_hi, dtype=np.uint64)
        lib().oracle_table_add_pairs(self.h, None if hi is None else _ptr(hi), _ptr(keys), _ptr(counts), len(keys))
        return self

    def lookup(self, key, canonicalise=False):
        """key: Python int (2 bits per base, first base most significant), any k <= 64."""
        key = int(key)
        return int(lib().oracle_lookup_wide(self.h, key >> 64, key & 0xFFFFFFFFFFFFFFFF, int(canonicalise)))

    def hist(self, low=1, high=10000, inc=1):
        b, c, n = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().oracle_hist_geometry(low, high, C.byref(b), C.byref(c), C.byref(n))
        bins = np.zeros(n.value, dtype=np.uint64)
        lib().oracle_hist(self.h, b.value, c.value, inc, n.value, _ptr(bins))
        return b.value, bins

    def gcp(self, cvg_scale=1.0, cvg_bins=1000):
        mx = np.zeros((self.k, cvg_bins + 1), dtype=np.uint64)
        lib().oracle_gcp(self.h, cvg_scale, cvg_bins, _ptr(mx))
        return mx

    def __del__(self):
        try:
            if self.h:
                lib().oracle_table_free(self.h)
                self.h = None
        except Exception:
            pass


def comp(t1, t2, t3=None, d1_scale=1.0, d2_scale=1.0, d1_bins=1001, d2_bins=1001):
    n = min(d1_bins, d2_bins)
    main = np.zeros((d1_bins, d2_bins), dtype=np.uint64)
    extra = [np.zeros((d1_bins, d2_bins), dtype=np.uint64) for _ in range(3)] if t3 else [None] * 3
    counters = np.zeros(13, dtype=np.uint64)
    spectra = np.zeros((4, n), dtype=np.uint64)
    lib().oracle_comp(t1.h, t2.h, t3.h if t3 else None, d1_scale, d2_scale, d1_bins, d2_bins,
                      _ptr(main), *[(_ptr(e) if e is not None else None) for e in extra],
                      _ptr(counters), _ptr(spectra))
    return {"main": main, "ends": extra[0], "middle": extra[1], "mixed": extra[2],
            "counters": dict(zip(CC_NAMES, (int(x) for x in counters))), "spectra": spectra}


def distances(s1, s2):
    s1 = np.ascontiguousarray(s1, dtype=np.uint64)
    s2 = np.ascontiguousarray(s2, dtype=np.uint64)
    out = np.zeros(5, dtype=np.float64)
    lib().oracle_distances(_ptr(s1), _ptr(s2), len(s1), _ptr(out))
    return out


def have_ref():
    return os.access(REF_KAT, os.X_OK)


def run_ref(args, cwd=None, timeout=3600):
    """Run the unmodified reference `kat` (oracle/_ref/kat)."""
    return subprocess.run([REF_KAT] + list(args), cwd=cwd, stdout=subprocess.PIPE,
                          stderr=subprocess.STDOUT, timeout=timeout, text=True)


def read_hist(path):
    """.hist body -> (labels, values) numpy arrays."""
    lab, val = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            a, b = line.split()
            lab.append(int(a)); val.append(int(b))
    return np.array(lab, dtype=np.uint64), np.array(val, dtype=np.uint64)


def read_mx(path):
    """.mx -> (header dict, matrix)."""
    hdr, rows = {}, []
    with open(path) as f:
        for line in f:
            if line.startswith("#"):
                if ":" in line:
                    k, v = line[1:].split(":", 1)
                    hdr[k.strip()] = v.strip()
                continue
            if line.strip():
                rows.append(np.array(line.split(), dtype=np.uint64))
    return hdr, np.vstack(rows) if rows else np.zeros((0, 0), dtype=np.uint64)


# ---- bench tooling beside the oracle (oracle/synth_ref.c): no product library involved ----
_synth = None


def synth_lib():
    global _synth
    if _synth is not None:
        return _synth
    so = os.path.join(ORACLE_DIR, "libsynthref.so")
    src = os.path.join(ORACLE_DIR, "synth_ref.c")
    if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libsynthref.so"], stdout=subprocess.DEVNULL)
    L = C.CDLL(so)
    L.synthref_reads.restype = C.c_int
    L.synthref_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32,
                                 C.c_double, C.c_double, C.c_int, C.c_void_p]
    L.synthref_instances.restype = C.c_uint64
    L.synthref_instances.argtypes = [C.c_void_p, C.c_uint64, C.c_int]
    L.synthref_write_fastq.restype = C.c_int
    L.synthref_write_fastq.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_char_p]
    _synth = L
    return L


def synth_reads(genome_len, seed_g, seed_r, first_read, n_reads, read_len=150, sub_rate=0.005, n_rate=0.0001, threads=None):
    """SURVEY.md §8(d) synthetic reads as a base stream (read,'N',...), generated on the CPU
    without libkatgpu; byte-equal to katgpu_synth_reads_host (tests/test_abi_and_host.py)."""
    out = np.empty(n_reads * (read_len + 1), dtype=np.uint8)
    rc = synth_lib().synthref_reads(genome_len, seed_g, seed_r, first_read, n_reads, read_len, sub_rate, n_rate,
                                    threads or (os.cpu_count() or 1), _ptr(out))
    if rc != 0:
        raise ValueError("synth_reads: bad parameters")
    return out


def count_instances(stream, k):
    a = np.frombuffer(stream, dtype=np.uint8) if not isinstance(stream, np.ndarray) else stream
    return int(synth_lib().synthref_instances(_ptr(a), a.size, k)) if a.size else 0


def write_fastq(stream, n_reads, read_len, path):
    if synth_lib().synthref_write_fastq(_ptr(stream), n_reads, read_len, os.fsencode(path)) != 0:
        raise OSError("write_fastq failed: " + str(path))

"""ctypes binding of include/katgpu.h.  Names mirror the reference classes the
C-ABI replaces: Context ~ process-wide device state, Table ~ InputHandler's
(HashCounter, LargeHashArrayPtr) pair (lib/include/kat/input_handler.hpp:53-55)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class KatGpuError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"katgpu status {status}: {msg}")
        self.status = status


def lib_path():
    return os.environ.get("KATGPU_LIB") or os.path.join(_HERE, "lib", "libkatgpu.so")


class TableInfo(C.Structure):
    _fields_ = [("k", C.c_int), ("canonical", C.c_int), ("slots", C.c_uint64), ("bytes", C.c_uint64),
                ("n_distinct", C.c_uint64), ("n_total", C.c_uint64), ("count_bits", C.c_uint32),
                ("grow_events", C.c_uint32), ("overflow_keys", C.c_uint64)]


CC_FIELDS = ["hash1_total", "hash2_total", "hash3_total", "hash1_distinct", "hash2_distinct", "hash3_distinct",
             "hash1_only_total", "hash2_only_total", "hash1_only_distinct", "hash2_only_distinct",
             "shared_hash1_total", "shared_hash2_total", "shared_distinct"]


class CompCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in CC_FIELDS]


# every symbol include/katgpu.h declares: (restype, argtypes)
_vp, _u64, _sz, _i, _d, _u32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_int, C.c_double, C.c_uint32
SIGNATURES = {
    "katgpu_last_error": (C.c_char_p, []),
    "katgpu_version": (_i, []),
    "katgpu_device_count": (_i, []),
    "katgpu_init": (_i, [_i, C.POINTER(_vp)]),
    "katgpu_shutdown": (None, [_vp]),
    "katgpu_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "katgpu_host_free": (None, [_vp]),
    "katgpu_count_begin": (_i, [_vp, _i, _i, _u64, _i, C.POINTER(_vp)]),
    "katgpu_count_feed": (_i, [_vp, _vp, _sz]),
    "katgpu_count_feed_device": (_i, [_vp, _vp, _sz]),
    "katgpu_count_feed_fastq": (_i, [_vp, _vp, _sz, C.c_uint, C.POINTER(_i)]),
    "katgpu_count_reserve": (_i, [_vp, _u64]),
    "katgpu_count_add_pairs": (_i, [_vp, _vp, _vp, _sz]),
    "katgpu_count_end": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64)]),
    "katgpu_table_free": (None, [_vp]),
    "katgpu_hist": (_i, [_vp, _u64, _u64, _u64, _u64, _vp]),
    "katgpu_gcp": (_i, [_vp, _d, _u32, _vp]),
    "katgpu_comp": (_i, [_vp, _vp, _vp, _d, _d, _u32, _u32, _vp, _vp, _vp, _vp, C.POINTER(CompCounters), _vp]),
    "katgpu_lookup": (_i, [_vp, _vp, _sz, _i, _vp]),
    "katgpu_table_export": (_i, [_vp, _vp, _vp, _sz, C.POINTER(_u64)]),
    "katgpu_table_export_filtered": (_i, [_vp, _u64, _u64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _sz, C.POINTER(_u64)]),
    "katgpu_lookup_wide": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "katgpu_table_export_wide": (_i, [_vp, _vp, _vp, _vp, _sz, C.POINTER(_u64)]),
    "katgpu_table_export_filtered_wide": (_i, [_vp, _u64, _u64, C.c_uint32, C.c_uint32, C.c_int, _vp, _vp, _vp, _sz, C.POINTER(_u64)]),
    "katgpu_count_add_pairs_wide": (_i, [_vp, _vp, _vp, _vp, _sz]),
    "katgpu_table_get_info": (_i, [_vp, C.POINTER(TableInfo)]),
    "katgpu_debug_set_count_bits": (None, [_i]),
    "katgpu_route_device": (_i, [_vp, _vp, _sz, _i, C.POINTER(_vp), _u64, C.POINTER(_u64)]),
    "katgpu_insert_hashed_device": (_i, [_vp, _vp, _u64]),
    "katgpu_device_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "katgpu_device_free": (None, [_vp, _vp]),
    "katgpu_ipc_export": (_i, [_vp, _vp]),
    "katgpu_ipc_import": (_i, [_vp, C.POINTER(_vp)]),
    "katgpu_ipc_close": (None, [_vp]),
    "katgpu_synchronize": (_i, [_vp]),
    "katgpu_release_cached_memory": (_i, [_vp]),
    "katgpu_dist_unique_id": (_i, [_vp]),
    "katgpu_dist_init": (_i, [_vp, _i, _i, _vp]),
    "katgpu_dist_rank": (_i, [_vp]),
    "katgpu_dist_n_ranks": (_i, [_vp]),
    "katgpu_dist_barrier": (_i, [_vp]),
    "katgpu_debug_route_local": (_i, [_vp, _vp, _sz, _i, C.POINTER(_d), C.POINTER(_u64)]),
    "katgpu_dist_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "katgpu_launch_count": (_u64, [_vp]),
    "katgpu_profile_enable": (_i, [_vp, _i]),
    "katgpu_profile_get": (_i, [_vp, C.c_char_p, C.POINTER(_d), C.POINTER(_u64)]),
    "katgpu_synth_reads_device": (_i, [_vp, _u64, _u64, _u64, _u64, _u64, _u32, _d, _d, _vp]),
    "katgpu_synth_reads_host": (_i, [_u64, _u64, _u64, _u64, _u64, _u32, _d, _d, _vp]),
    "katgpu_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "katgpu_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
}


def lib():
    """Load libkatgpu.so (in-tree).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise KatGpuError(-100, f"{p} is missing: build it with `make` (python __graft_entry__.py build); "
                                    "there is no fallback implementation")
        L = C.CDLL(p)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _check(st):
    if st != 0:
        raise KatGpuError(st, lib().katgpu_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def hist_geometry(low=1, high=10000):
    """base / ceil / nb_buckets of src/histogram.hpp:172-178, src/histogram.cc:68-70."""
    base = (1 if 1 >= low else low - 1) if low > 1 else 1
    ceil = high + 1
    return base, ceil, ceil + 1 - base


class Context:
    def __init__(self, device=0):
        import weakref
        self.h = C.c_void_p()
        self._tables = weakref.WeakSet()
        _check(lib().katgpu_init(device, C.byref(self.h)))

    def close(self):
        if self.h:
            for t in list(self._tables):  # tables must not outlive their context
                t.free()
            lib().katgpu_shutdown(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # --- counting -------------------------------------------------------------
    def count_begin(self, k, canonical=True, hash_size=100_000_000, allow_grow=True):
        t = C.c_void_p()
        _check(lib().katgpu_count_begin(self.h, k, int(canonical), int(hash_size), int(allow_grow), C.byref(t)))
        return Table(self, t, k, canonical)

    def count(self, chunks, k, canonical=True, hash_size=100_000_000, allow_grow=True):
        """InputHandler::count equivalent: one feed per chunk (no k-mer spans chunks)."""
        t = self.count_begin(k, canonical, hash_size, allow_grow)
        for c in chunks:
            t.feed(c)
        t.end()
        return t

    # --- memory / plumbing -----------------------------------------------------
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        _check(lib().katgpu_device_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, p):
        lib().katgpu_device_free(self.h, p)

    def h2d(self, dst, src_np):
        _check(lib().katgpu_memcpy_h2d(self.h, dst, _ptr(src_np), src_np.nbytes))

    def d2h(self, dst_np, src):
        _check(lib().katgpu_memcpy_d2h(self.h, _ptr(dst_np), src, dst_np.nbytes))

    def synchronize(self):
        _check(lib().katgpu_synchronize(self.h))

    def release_cached_memory(self):
        """Give the device blocks cached for reuse back to the driver (katgpu_release_cached_memory)."""
        _check(lib().katgpu_release_cached_memory(self.h))

    # --- multi-GPU job (one context per GPU) ------------------------------------
    def dist_init(self, rank, n_ranks, unique_id: bytes):
        """Join a job of n_ranks contexts; unique_id = dist_unique_id() of rank 0, handed to everybody."""
        assert len(unique_id) == 128
        buf = C.create_string_buffer(bytes(unique_id), 128)
        _check(lib().katgpu_dist_init(self.h, rank, n_ranks, buf))

    @property
    def rank(self):
        return int(lib().katgpu_dist_rank(self.h))

    @property
    def n_ranks(self):
        return int(lib().katgpu_dist_n_ranks(self.h))

    def dist_barrier(self):
        _check(lib().katgpu_dist_barrier(self.h))

    def dist_stats(self):
        a, b, r = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(lib().katgpu_dist_stats(self.h, C.byref(a), C.byref(b), C.byref(r)))
        return {"starts_routed": a.value, "keys_received": b.value, "rounds": r.value}

    def launch_count(self):
        return int(lib().katgpu_launch_count(self.h))

    def profile_enable(self, on=True):
        _check(lib().katgpu_profile_enable(self.h, int(on)))

    def profile_get(self, family):
        ms, n = C.c_double(), C.c_uint64()
        _check(lib().katgpu_profile_get(self.h, family.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def synth_reads_device(self, out_dev, genome_len, seed_g, seed_r, first_read, n_reads, read_len=150,
                           sub_rate=0.005, n_rate=0.0001):
        _check(lib().katgpu_synth_reads_device(self.h, genome_len, seed_g, seed_r, first_read, n_reads, read_len,
                                               sub_rate, n_rate, out_dev))


def dist_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().katgpu_dist_unique_id(buf))
    return buf.raw


def synth_reads_host(genome_len, seed_g, seed_r, first_read, n_reads, read_len=150, sub_rate=0.005, n_rate=0.0001):
    out = np.empty(n_reads * (read_len + 1), dtype=np.uint8)
    _check(lib().katgpu_synth_reads_host(genome_len, seed_g, seed_r, first_read, n_reads, read_len, sub_rate,
                                         n_rate, _ptr(out)))
    return out


class CompResult(dict):
    pass


class Table:
    def __init__(self, ctx, h, k, canonical):
        self.ctx, self.h, self.k, self.canonical = ctx, h, k, bool(canonical)
        self.n_distinct = self.n_total = None
        ctx._tables.add(self)

    def feed(self, bases):
        """bases: bytes / numpy uint8 array in HOST memory."""
        a = np.frombuffer(bases, dtype=np.uint8) if isinstance(bases, (bytes, bytearray, memoryview)) else bases
        if a.size:
            _check(lib().katgpu_count_feed(self.h, _ptr(a), a.size))
        return self

    def feed_fastq(self, raw, trim5p=0):
        """raw: bytes / numpy uint8 of WHOLE 4-line FASTQ records in host memory; parsed on the GPU.
        Returns True when the bytes were regular and have been counted, False when nothing was counted."""
        a = np.frombuffer(raw, dtype=np.uint8) if isinstance(raw, (bytes, bytearray, memoryview)) else raw
        reg = C.c_int(0)
        _check(lib().katgpu_count_feed_fastq(self.h, _ptr(a) if a.size else None, a.size, trim5p, C.byref(reg)))
        return bool(reg.value)

    def feed_ptr(self, host_ptr, n):
        _check(lib().katgpu_count_feed(self.h, host_ptr, n))
        return self

    def feed_device(self, dev_ptr, n):
        _check(lib().katgpu_count_feed_device(self.h, dev_ptr, n))
        return self

    def add_pairs(self, keys, counts):
        """(k-mer, count) records in HOST memory: count[keys[i]] += counts[i] (HashLoader::loadHash)."""
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        assert k.size == c.size
        if k.size:
            _check(lib().katgpu_count_add_pairs(self.h, _ptr(k), _ptr(c), k.size))
        return self

    def add_pairs_wide(self, keys_hi, keys_lo, counts):
        """add_pairs with two words per key (any k; keys_hi is ignored for k <= 32)."""
        hi = np.ascontiguousarray(keys_hi, dtype=np.uint64)
        lo = np.ascontiguousarray(keys_lo, dtype=np.uint64)
        c = np.ascontiguousarray(counts, dtype=np.uint64)
        assert hi.size == lo.size == c.size
        if lo.size:
            _check(lib().katgpu_count_add_pairs_wide(self.h, _ptr(hi), _ptr(lo), _ptr(c), lo.size))
        return self

    def lookup_wide(self, keys_hi, keys_lo, canonicalise=False):
        """Counts of two-word keys (any k): keys_lo = last 32 bases, keys_hi = the first k-32."""
        hi = np.ascontiguousarray(keys_hi, dtype=np.uint64)
        lo = np.ascontiguousarray(keys_lo, dtype=np.uint64)
        out = np.zeros(lo.size, dtype=np.uint64)
        _check(lib().katgpu_lookup_wide(self.h, _ptr(hi), _ptr(lo), lo.size, int(canonicalise), _ptr(out)))
        return out

    def export_wide(self, bounds=None, inside=True):
        """(keys_hi, keys_lo, counts, n) sorted by key; bounds = (low_count, high_count, low_gc, high_gc) selects the
        k-mers inside (or outside) them as `kat filter kmer` does."""
        n = C.c_uint64()

        def call(hi, lo, cn, cap):
            if bounds is None:
                return lib().katgpu_table_export_wide(self.h, hi, lo, cn, cap, C.byref(n))
            return lib().katgpu_table_export_filtered_wide(self.h, *[int(b) for b in bounds], int(bool(inside)), hi, lo, cn, cap, C.byref(n))

        _check(call(None, None, None, 0))
        cap = n.value
        hi, lo, cn = (np.zeros(max(cap, 1), dtype=np.uint64) for _ in range(3))
        _check(call(_ptr(hi), _ptr(lo), _ptr(cn), cap))
        assert n.value == cap
        order = np.lexsort((lo[:cap], hi[:cap]))
        return hi[:cap][order], lo[:cap][order], cn[:cap][order], cap

    def reserve(self, expected_kmers):
        _check(lib().katgpu_count_reserve(self.h, int(expected_kmers)))
        return self

    def end(self):
        d, t = C.c_uint64(), C.c_uint64()
        _check(lib().katgpu_count_end(self.h, C.byref(d), C.byref(t)))
        self.n_distinct, self.n_total = d.value, t.value
        return self

    def info(self):
        i = TableInfo()
        _check(lib().katgpu_table_get_info(self.h, C.byref(i)))
        return {n: getattr(i, n) for n, _ in TableInfo._fields_}

    def hist(self, low=1, high=10000, inc=1):
        base, ceil, nb = hist_geometry(low, high)
        bins = np.zeros(nb, dtype=np.uint64)
        _check(lib().katgpu_hist(self.h, base, ceil, inc, nb, _ptr(bins)))
        return base, bins

    def gcp(self, cvg_scale=1.0, cvg_bins=1000):
        mx = np.zeros((self.k, cvg_bins + 1), dtype=np.uint64)
        _check(lib().katgpu_gcp(self.h, cvg_scale, cvg_bins, _ptr(mx)))
        return mx

    def lookup(self, keys, canonicalise=False):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.zeros(keys.size, dtype=np.uint64)
        _check(lib().katgpu_lookup(self.h, _ptr(keys), keys.size, int(canonicalise), _ptr(out)))
        return out

    def export(self, sort=True):
        n = C.c_uint64()
        cap = int(self.n_distinct if self.n_distinct is not None else self.info()["n_distinct"])
        keys = np.zeros(max(cap, 1), dtype=np.uint64)
        counts = np.zeros(max(cap, 1), dtype=np.uint64)
        _check(lib().katgpu_table_export(self.h, _ptr(keys), _ptr(counts), cap, C.byref(n)))
        m = min(cap, n.value)
        if not sort:
            return keys[:m], counts[:m], n.value
        order = np.argsort(keys[:m], kind="stable")
        return keys[:m][order], counts[:m][order], n.value

    def export_filtered(self, low_count, high_count, low_gc, high_gc, inside=True):
        """The k-mers inside (or outside) the count / GC bounds of `kat filter kmer`, sorted by key."""
        n = C.c_uint64()
        args = (self.h, int(low_count), int(high_count), int(low_gc), int(high_gc), int(bool(inside)))
        _check(lib().katgpu_table_export_filtered(*args, None, None, 0, C.byref(n)))
        cap = n.value
        keys = np.zeros(max(cap, 1), dtype=np.uint64)
        counts = np.zeros(max(cap, 1), dtype=np.uint64)
        _check(lib().katgpu_table_export_filtered(*args, _ptr(keys), _ptr(counts), cap, C.byref(n)))
        assert n.value == cap
        order = np.argsort(keys[:cap], kind="stable")
        return keys[:cap][order], counts[:cap][order]

    def route_device(self, bases_dev, n, out_ptrs, cap_keys):
        n_ranks = len(out_ptrs)
        arr = (C.c_void_p * n_ranks)(*out_ptrs)
        counts = (C.c_uint64 * n_ranks)()
        _check(lib().katgpu_route_device(self.h, bases_dev, n, n_ranks, arr, cap_keys, counts))
        return [int(x) for x in counts]

    def insert_hashed_device(self, dev_ptr, n):
        _check(lib().katgpu_insert_hashed_device(self.h, dev_ptr, n))

    def free(self):
        if self.h and self.ctx.h:
            lib().katgpu_table_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def comp(t1, t2, t3=None, d1_scale=1.0, d2_scale=1.0, d1_bins=1001, d2_bins=1001):
    n = min(d1_bins, d2_bins)
    main = np.zeros((d1_bins, d2_bins), dtype=np.uint64)
    extra = [np.zeros((d1_bins, d2_bins), dtype=np.uint64) for _ in range(3)] if t3 else [None] * 3
    cc = CompCounters()
    spectra = np.zeros((4, n), dtype=np.uint64)
    _check(lib().katgpu_comp(t1.h, t2.h, t3.h if t3 else None, d1_scale, d2_scale, d1_bins, d2_bins, _ptr(main),
                             *[(_ptr(e) if e is not None else None) for e in extra], C.byref(cc), _ptr(spectra)))
    return CompResult(main=main, ends=extra[0], middle=extra[1], mixed=extra[2],
                      counters={f: int(getattr(cc, f)) for f in CC_FIELDS}, spectra=spectra)

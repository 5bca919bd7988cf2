#!/usr/bin/env python
"""Two-word k-mers (k = 33..64) on one GPU: kernel time of the extract + insert kernel and of the walks for a synthetic
read set (SURVEY.md §8(d) generator), next to the same job at k = 27 through the default one-word path.
Usage: wide_probe.py [reads] [k ...]     (default: 10 M reads of 150 bp = 1.5 Gbp; k = 27 33 41 64)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kat_b200

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ks = [int(x) for x in sys.argv[2:]] or [27, 33, 41, 64]
with kat_b200.Context(0) as ctx:
    n = reads * 151
    d = ctx.device_alloc(n + 64)
    ctx.synth_reads_device(d, 200_000_000, 1, 2, 0, reads, 150, 0.005, 0.0001)
    for k in ks:
        for rep in range(2):  # the second run reuses the cached device blocks
            ctx.profile_enable(True)
            t0 = time.perf_counter()
            t = ctx.count_begin(k, True, 1 << 29, True)
            t.feed_device(d, n).end()
            ctx.synchronize()
            wall = time.perf_counter() - t0
            base, bins = t.hist()
            mx = t.gcp()
            fam = {f: ctx.profile_get(f) for f in ("count", "bin1", "bin2", "tile", "insert", "rehash", "memset", "hist", "gcp")}
            ctx.profile_enable(False)
            info = t.info()
            kern = sum(fam[f][0] for f in ("count", "bin1", "bin2", "tile", "insert", "rehash"))
            print(f"k={k} run {rep}: {t.n_total} k-mers, {t.n_distinct} distinct, slots 2^{info['slots'].bit_length() - 1}, grow {info['grow_events']}; "
                  f"counting kernels {kern:.1f} ms = {t.n_total / kern / 1e6:.2f} G k-mers/s (wall {wall * 1e3:.0f} ms); "
                  f"hist {fam['hist'][0]:.2f} ms, gcp {fam['gcp'][0]:.2f} ms; "
                  + ", ".join(f"{f} {v[0]:.1f}" for f, v in fam.items() if v[0] and f not in ("hist", "gcp")))
            assert int(bins.sum()) == t.n_distinct and int(mx.sum()) <= t.n_distinct
            t.free()
    ctx.device_free(d)

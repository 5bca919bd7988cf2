#!/usr/bin/env python
"""route_bases_kernel alone, one GPU: time per 10 Gbp shard with 1/2/4/8 owners whose receive areas are
local memory (no NVLink) — what the multi-GPU extraction costs before the wire.  Usage: route_probe.py [reads]"""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kat_b200
from kat_b200 import binding as B

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 66_666_667
with kat_b200.Context(0) as ctx:
    n = reads * 151
    d = ctx.device_alloc(n + 64)
    ctx.synth_reads_device(d, 200_000_000, 1, 2, 0, reads, 150, 0.005, 0.0001)
    t = ctx.count_begin(27, True, 1 << 20, True)
    chunk = 128 << 20
    for nr in (1, 2, 4, 8):
        tot_ms, tot_keys, off = 0.0, 0, 0
        while off < n:  # the same 128 MiB rounds as the job
            m = min(chunk, n - off)
            ms, keys = C.c_double(), C.c_uint64()
            B._check(B.lib().katgpu_debug_route_local(t.h, d + off, m, nr, C.byref(ms), C.byref(keys)))
            tot_ms += ms.value; tot_keys += keys.value; off += m
        print(f"owners {nr}: {tot_ms:8.2f} ms for {tot_keys} keys ({tot_keys / tot_ms / 1e6:.1f} G keys/s, {tot_keys * 8 / tot_ms / 1e6:.0f} GB/s of keys written locally)")
    t.end(); t.free(); ctx.device_free(d)

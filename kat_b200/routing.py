"""Host-side restatement of the routing arithmetic of table.cuh (mix_key,
owner_of) — used by the multi-GPU orchestration tests; the product computes
these on the GPU."""
import numpy as np

KMERS_SLACK = 1 << 16
_M1 = np.uint64(0xff51afd7ed558ccd)
_M2 = np.uint64(0xc4ceb9fe1a85ec53)
_OWN = np.uint64(0x9E3779B97F4A7C15)


def mix_key(x, k):
    """kat_b200/csrc/table.cuh: mix_key — a bijection on 2k-bit integers."""
    b = 2 * k
    km = np.uint64((1 << b) - 1) if b < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    sh = np.uint64((b + 1) // 2)
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> sh)
        x = (x * _M1) & km
    return x


def owner_of(h, world):
    """kat_b200/csrc/table.cuh: owner_of."""
    with np.errstate(over="ignore"):
        y = (np.asarray(h, dtype=np.uint64) * _OWN) >> np.uint64(32)
        return ((y * np.uint64(world)) >> np.uint64(32)).astype(np.int64)

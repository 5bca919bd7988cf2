"""kat_b200 — Python binding of libkatgpu, the B200-native k-mer spectra engine
behind `kat hist` / `kat gcp` / `kat comp`.

The product is the CUDA library (kat_b200/csrc -> kat_b200/lib/libkatgpu.so)
and the C++ `kat` front-end (kat_b200/bin/kat); this package is the thin
ctypes layer used by tests, bench.py and the multi-GPU driver.  There is no CPU
fallback: importing works anywhere (so the ABI can be inspected), but creating
a Context without a CUDA device raises KatGpuError.
"""
from .binding import (  # noqa: F401
    CompResult, Context, KatGpuError, Table, hist_geometry, lib, lib_path,
)

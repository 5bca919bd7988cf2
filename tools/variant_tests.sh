#!/bin/bash
# Parity tests of a tuning variant library: tools/variant_tests.sh <name> [pytest -k expression]
KATGPU_LIB=$PWD/kat_b200/lib/variants/libkatgpu_$1.so timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "${2:-tiled}" 2>&1 | tail -4

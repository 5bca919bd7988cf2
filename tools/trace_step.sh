#!/bin/bash
# Host-side phase times of one resident bench step (KATGPU_TRACE) — where a step's wall time goes beyond its kernels.
KATGPU_TRACE=1 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline "$@" 2>&1 | grep -v "sweep_tiled:" | tail -45 | cut -c1-300

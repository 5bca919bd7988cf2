#!/bin/bash
# Host-side phase times of one resident bench step (KATGPU_TRACE) — where a step's wall time goes beyond its kernels.
KATGPU_TRACE=1 python bench.py --steps ${STEPS:-4} --warmup 1 --no-e2e --no-cpu-baseline "$@" 2>&1 | grep -v "sweep_tiled:\|alloc_table_storage\|pend_alloc \|table_free\|katgpu_init\|katgpu_count_begin" | tail -${LINES:-70} | cut -c1-300

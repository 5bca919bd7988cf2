#!/bin/bash
# A new kernel goes through a short, time-boxed smoke first; the long steps only run when it passes
# (a hung kernel must not eat the GPU lease): tools/safe_check.sh <outdir> <gpu.sh steps...>
out=$1; shift
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "smoke failed or timed out: stopping"; exit 1; fi
KATGPU_COUNT_PATH=tiled timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "larger_tables or rare_paths" 2>&1 | tail -2
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "tiled-path tests failed or timed out: stopping"; exit 1; fi
exec tools/gpu.sh $out "$@"

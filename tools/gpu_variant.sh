#!/bin/bash
# Parity tests (ctypes path honours KATGPU_LIB) and a bench line for each variant library.
out=gpurun_out/${1:-variant}; shift; mkdir -p $out
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; e=d.get("e2e") or {}; print("value %.2f G/s ms/step %.1f  %s other %s launches %d e2e %.2f G/s %.1f ms" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}, d["gpu_launches"], (e.get("value") or 0)/1e9, e.get("ms_per_step") or 0))'
for v in "$@"; do
  lib=$PWD/kat_b200/lib/variants/libkatgpu_$v.so
  echo -n "$v bench: "; KATGPU_LIB=$lib timeout 400 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>$out/$v.err > $out/$v.json; python -c "$P" < $out/$v.json 2>/dev/null || tail -5 $out/$v.err
  (KATGPU_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=5 -k "tiled") > $out/tests_$v.log 2>&1; echo -n "$v tests: "; grep -E "passed|failed" $out/tests_$v.log | tail -1
done

#!/bin/bash
# One gpurun call: GPU parity tests on the default library, then A/B benches of library variants
# (KATGPU_LIB) and env-var shapes.  Output: gpurun_out/$1/
out=gpurun_out/${1:-ab}; mkdir -p $out
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; e=d.get("e2e") or {}; print("value %.2f G/s ms/step %.1f  %s other %s launches %d e2e %.2f G/s" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}, d["gpu_launches"], (e.get("value") or 0)/1e9))'
(time python -m pytest tests -m gpu -x -q) > $out/tests.log 2>&1; tail -4 $out/tests.log
run() { name=$1; shift; echo -n "$name: "; timeout 400 env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline $BENCHFLAGS 2>$out/$name.err > $out/$name.json; python -c "$P" < $out/$name.json || tail -3 $out/$name.err; }
BENCHFLAGS="" run new_full A=1
BENCHFLAGS="--no-e2e"
run base KATGPU_LIB=$PWD/kat_b200/lib/variants/libkatgpu_base.so
run new_noe2e A=1
for v in u2c4 u8c2 u4t128c6 u4c4; do run $v KATGPU_LIB=$PWD/kat_b200/lib/variants/libkatgpu_$v.so; done
run pb10 KATGPU_MAX_PB=10
run pb11_16mb KATGPU_MAX_PB=11 KATGPU_SLICE_MB=16
run nopf KATGPU_SWEEP_PREFETCH=0
run slow_sweep KATGPU_SWEEP_FAST=0
BENCHFLAGS="--no-e2e --slots-log2 31" run slots31 A=1
BENCHFLAGS="--no-e2e --slots-log2 31" run slots31_pb10 KATGPU_MAX_PB=10 KATGPU_SLICE_MB=16

#!/bin/bash
out=gpurun_out/${1:-quick}; mkdir -p $out
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; e=d.get("e2e") or {}; print("value %.2f G/s ms/step %.1f  %s other %s launches %d e2e %.2f G/s %.1f ms" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}, d["gpu_launches"], (e.get("value") or 0)/1e9, e.get("ms_per_step") or 0))'
(time timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "tiled or larger") > $out/tests_tiled.log 2>&1; grep -E "passed|failed|^FAILED|^ERROR" $out/tests_tiled.log | head -30
run() { name=$1; shift; echo -n "$name: "; timeout 500 env "$@" python bench.py --steps 3 --warmup 2 --no-cpu-baseline $BENCHFLAGS 2>$out/$name.err > $out/$name.json; python -c "$P" < $out/$name.json 2>/dev/null || tail -5 $out/$name.err; }
BENCHFLAGS="" run default A=1

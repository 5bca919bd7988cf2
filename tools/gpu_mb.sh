#!/bin/bash
out=gpurun_out/${1:-mb}; mkdir -p $out
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("value %.2f G/s ms/step %.1f  %s other %s launches %d" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}, d["gpu_launches"]))'
./tools/microbench | tee $out/microbench.txt
run() { name=$1; shift; echo -n "$name: "; timeout 400 env "$@" python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e $BENCHFLAGS 2>$out/$name.err > $out/$name.json; python -c "$P" < $out/$name.json || tail -3 $out/$name.err; }
run fast_default A=1
run loads_only KATGPU_SWEEP_EXPERIMENT=1
run reds_only KATGPU_SWEEP_EXPERIMENT=2
(time python -m pytest tests -m gpu -x -q) > $out/tests.log 2>&1; grep -E "passed|failed" $out/tests.log

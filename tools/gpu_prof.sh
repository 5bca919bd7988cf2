#!/bin/bash
# One gpurun call: ncu --set full of the timed step's partition + sweep kernels at 5 Gbp / 2^31 slots
# (the 10 Gbp job does not fit ncu's save/restore), then timing-only experiments at full size.
out=gpurun_out/${1:-prof}; mkdir -p $out
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("value %.2f G/s ms/step %.1f  %s other %s launches %d" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}, d["gpu_launches"]))'
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:sweep_kernel|partition_kernel' -s 2 -c 2 -o $out/prof \
   python bench.py --reads 33000000 --slots-log2 31 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/ncu.log 2>&1; tail -3 $out/ncu.log
run() { name=$1; shift; echo -n "$name: "; timeout 400 env "$@" python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e $BENCHFLAGS 2>$out/$name.err > $out/$name.json; python -c "$P" < $out/$name.json || tail -3 $out/$name.err; }
run loads_only KATGPU_SWEEP_EXPERIMENT=1
run reds_only KATGPU_SWEEP_EXPERIMENT=2
run window0 KATGPU_SWEEP_WINDOW=0
run window2 KATGPU_SWEEP_WINDOW=2
BENCHFLAGS="--reads 33000000 --slots-log2 31" run half_new A=1
BENCHFLAGS="--reads 33000000 --slots-log2 31" run half_base KATGPU_LIB=$PWD/kat_b200/lib/variants/libkatgpu_base.so
nvidia-smi -q | grep -i -E "ecc|Current|Product Name|Power Limit|Default Power" | head -20

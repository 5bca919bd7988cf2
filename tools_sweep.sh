#!/bin/bash
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("value %.2f G/s ms/step %.1f  %s other %s launches %d" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}, d["gpu_launches"]))'
for L in 31 32; do
  echo -n "slots 2^$L 5Gbp: "; KATGPU_COUNT_PATH=partition timeout 200 python bench.py --reads 33000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --slots-log2 $L 2>/dev/null | python -c "$P"
done
for L in 31 32; do
echo -n "full 2^$L: "; timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --slots-log2 $L 2>gpurun_out/b_full6_$L.err | tee gpurun_out/b_full6_$L.json | python -c "$P"; tail -2 gpurun_out/b_full6_$L.err
python -c "
import json; d=json.load(open('gpurun_out/b_full6_$L.json')); print('e2e', d['e2e']['value']/1e9, d['e2e']['ms_per_step'])"
done

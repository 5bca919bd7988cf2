#!/bin/bash
# tuning sweep (run on the GPU box): sweep-kernel shape variants of the partitioned path (256-bit bucket loads)
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("value %.2f G/s  %s  launches %d" % (d["value"]/1e9, {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, d["gpu_launches"]))'
for v in default t256c4u2 t256c4u1 t512c2u2 t256c4u8; do
    lib=$PWD/kat_b200/lib/variants/libkatgpu_$v.so; [ $v = default ] && lib=$PWD/kat_b200/lib/libkatgpu.so
    sp=2; [ $v = t512c2u2 ] && sp=1
    echo -n "$v split=$sp : "
    KATGPU_LIB=$lib KATGPU_COUNT_PATH=partition KATGPU_SWEEP_SPLIT=$sp timeout 200 python bench.py --reads 8000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --slots-log2 30 2>/dev/null | python -c "$P"
done
echo -n "direct path (256-bit loads): "
KATGPU_COUNT_PATH=direct timeout 200 python bench.py --reads 8000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --slots-log2 30 2>/dev/null | python -c "$P"

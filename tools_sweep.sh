#!/bin/bash
# tuning sweep (run on the GPU box): variant x L2 fetch granularity
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("value %.2f G/s count_ms %.2f launches %d kmers %d" % (d["value"]/1e9, r["kernel_ms_total"]/d["steps"], r["kernel_launches"], d["kmer_instances_per_step"]))'
for v in u4c3 u4c4 u8c2 u2c4 u8c3; do
  for f in 32 128; do
    echo -n "$v fetch=$f : "
    KATGPU_LIB=$PWD/kat_b200/lib/variants/libkatgpu_$v.so KATGPU_L2_FETCH_BYTES=$f timeout 200 python bench.py --reads 4000000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --slots-log2 30 2>/dev/null | python -c "$P"
  done
done

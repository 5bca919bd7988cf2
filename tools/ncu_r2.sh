#!/bin/bash
# ncu evidence of round 2 (one GPU): launch list of a full-size step, then --set full of the K1 kernels and the walks
# at a size ncu's kernel replay can hold (5 Gbp, 2^30 slots).  Reports land in gpurun_out/$1/.
out=gpurun_out/${1:-ncu}; mkdir -p $out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/launches.log 2>&1; tail -1 $out/launches.log | cut -c1-200; wc -l $out/launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:bin_bases_kernel|bin_keys_kernel|tile_insert_kernel' -c 4 -o $out/k1 python bench.py --reads 33000000 --slots-log2 30 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $out/k1.log 2>&1; tail -2 $out/k1.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:hist_kernel|gcp_kernel|comp_kernel' -c 3 -o $out/walks python bench.py --workload comp --reads 20000000 --slots-log2 29 --steps 1 --warmup 0 --no-cpu-baseline > $out/walks.log 2>&1; tail -2 $out/walks.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:route_bases_kernel' -c 1 -o $out/route python tools/route_probe.py 6000000 > $out/route.log 2>&1; tail -2 $out/route.log | cut -c1-200
ls -la $out

#!/bin/bash
# ncu --set full of the timed step's bin_bases, first bin_keys and first tile_insert launches at
# 5 Gbp / 2^31 slots (one warm-up step = 1 + 2 x groups matching launches are skipped).
out=gpurun_out/${1:-prof2}; mkdir -p $out
SKIP=${2:-11}
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:bin_bases|bin_keys|tile_insert' -s $SKIP -c 3 -o $out/prof \
   python bench.py --reads 33000000 --slots-log2 31 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $out/ncu.log 2>&1; tail -2 $out/ncu.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $out/launches_bench.log 2>&1; tail -1 $out/launches_bench.log | cut -c1-200

#!/bin/bash
# The one script every gpurun call of this repo goes through (run from the repo root on the GPU box):
#   tools/gpu.sh <outdir-name> <step> [<step> ...]
# steps:
#   tests[:<pytest -k expr>]   pytest -m gpu (all, or a selection)
#   smoke                      __graft_entry__.smoke()
#   bench[:<extra flags>]      python bench.py <flags>             -> bench*.json  (flags: '+' separated, e.g. bench:--steps+5+--slots-log2+31)
#   ref[:<extra flags>]        python bench.py --impl reference    -> ref*.json
#   variant:<name>[:<flags>]   bench line with kat_b200/lib/variants/libkatgpu_<name>.so (tools/variants.sh builds them)
#   launches[:<flags>]         ncu launch list (gpu__time_duration) of a short bench run -> launches.csv
#   ncu:<kernel regex>[:<flags>]  ncu --set full of the matching kernels of a short, smaller bench run -> <regex>.ncu-rep
#   mb[:<args>]                tools/microbench (memory-system microbenchmarks)
#   torchrun:<N>[:<flags>]     bench.py under torch.distributed.run on N GPUs
#   sh:<command>               anything else ('+' = space)
out=gpurun_out/$1; shift; mkdir -p $out
P='import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1])
except Exception as e:
    print("no json line:", e); sys.exit(0)
if "unavailable" in d: print("unavailable:", d["unavailable"]); sys.exit(0)
r=d.get("roofline") or {}; e=d.get("e2e") or {}
print("value %.2f G/s  %.1f ms/step  frac %s  fam %s  other %s  launches %s  e2e %.2f G/s %.1f ms" % (d["value"]/1e9, d["ms_per_step"], r.get("frac") and round(r["frac"],4),
      {k: round(v/d["steps"],2) for k,v in (r.get("kernel_ms_by_family") or {}).items()}, {k: round(v/d["steps"],2) for k,v in (r.get("other_kernels_ms") or {}).items() if v},
      d.get("gpu_launches"), (e.get("value") or 0)/1e9, e.get("ms_per_step") or 0))
for k in ("cpu_baseline","cli_file","reduction","nvlink"):
    if d.get(k): print("  ", k, json.dumps(d[k])[:600])'
i=0
for step in "$@"; do
  i=$((i+1)); kind=${step%%:*}; rest=""; [ "$step" != "$kind" ] && rest=${step#*:}
  flags=$(echo "$rest" | tr '+' ' ')
  echo "=== [$i] $step"
  case $kind in
    tests) if [ -n "$rest" ]; then (time timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -k "$flags") > $out/tests_$i.log 2>&1; else (time timeout 2400 python -m pytest tests -m gpu -q --maxfail=10) > $out/tests_$i.log 2>&1; fi
           grep -E "passed|failed|^FAILED|^ERROR|^real" $out/tests_$i.log | head -30 ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    bench) timeout 1500 python bench.py $flags > $out/bench_$i.json 2> $out/bench_$i.err; python -c "$P" < $out/bench_$i.json; tail -2 $out/bench_$i.err ;;
    ref)   timeout 1500 python bench.py --impl reference $flags > $out/ref_$i.json 2> $out/ref_$i.err; cut -c1-900 $out/ref_$i.json; tail -2 $out/ref_$i.err ;;
    variant) name=${rest%%:*}; vf=""; [ "$rest" != "$name" ] && vf=$(echo "${rest#*:}" | tr '+' ' ')
           KATGPU_LIB=$PWD/kat_b200/lib/variants/libkatgpu_$name.so timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e $vf > $out/variant_$name.json 2> $out/variant_$name.err
           echo -n "$name: "; python -c "$P" < $out/variant_$name.json; tail -2 $out/variant_$name.err ;;
    launches) timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e $flags > $out/launches.log 2>&1; tail -2 $out/launches.log; wc -l $out/launches.csv ;;
    ncu)   rx=${rest%%:*}; nf=""; [ "$rest" != "$rx" ] && nf=$(echo "${rest#*:}" | tr '+' ' ')
           timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$rx" -c 6 -o $out/ncu_$i python bench.py --reads 33000000 --slots-log2 31 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e $nf > $out/ncu_$i.log 2>&1; tail -3 $out/ncu_$i.log; ls -la $out/ncu_$i.ncu-rep ;;
    mb)    timeout 900 tools/microbench $flags 2>&1 | tee $out/microbench_$i.txt | tail -60 ;;
    torchrun) n=${rest%%:*}; tf=""; [ "$rest" != "$n" ] && tf=$(echo "${rest#*:}" | tr '+' ' ')
           timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $tf > $out/bench_n${n}_$i.json 2> $out/bench_n${n}_$i.err
           python -c "$P" < $out/bench_n${n}_$i.json; tail -4 $out/bench_n${n}_$i.err ;;
    sh)    bash -c "$flags" 2>&1 | tail -40 ;;
    *) echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv,noheader | head -8

#!/bin/bash
# gpurun --gpus N: NCCL test + the scaling bench at N ranks
N=${2:-2}; out=gpurun_out/${1:-multi}; mkdir -p $out
(time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q) > $out/tests_multi.log 2>&1; grep -E "passed|failed|skipped" $out/tests_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > $out/bench_n$N.json 2> $out/bench_n$N.err
P='import json,sys; d=json.loads(sys.stdin.read()); r=d["roofline"]; print("n_gpus %d value %.2f G/s ms/step %.1f  %s other %s" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],2) for k,v in r["kernel_ms_by_family"].items()}, {k: round(v/d["steps"],2) for k,v in r["other_kernels_ms"].items() if v}))'
python -c "$P" < $out/bench_n$N.json || tail -20 $out/bench_n$N.err

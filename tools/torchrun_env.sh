#!/bin/bash
# bench.py under torchrun with extra environment: tools/torchrun_env.sh <N> "VAR=value ..." [bench flags]
n=$1; envs=$2; shift 2
env $envs timeout ${TIMEOUT:-300} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n "$@" 2>/tmp/torchrun_env.err | python -c '
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; e=d.get("e2e") or {}
print("value %.2f G/s  %.1f ms/step  fam %s  e2e %.2f G/s" % (d["value"]/1e9, d["ms_per_step"], {k: round(v/d["steps"],1) for k,v in r["kernel_ms_by_family"].items()}, (e.get("value") or 0)/1e9))'
tail -3 /tmp/torchrun_env.err | cut -c1-300

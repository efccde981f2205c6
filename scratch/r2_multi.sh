#!/bin/bash
# usage: r2_multi.sh N   -> gpurun_out/r2_n${N}_${config}.json
N=$1
mkdir -p gpurun_out
run() { c=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2970$RANDOM_SUFFIX bench.py --gpus $N --config $c "$@" > gpurun_out/r2_n${N}_$c.json 2> gpurun_out/r2_n${N}_$c.err
  echo "$c rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_n${N}_$c.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","warmup")}, d.get("e2e",{}).get("value"), d.get("value_hot_path",{}).get("value"), d.get("prune"))
except Exception as e:
    print("no json", e); print(open('gpurun_out/r2_n${N}_$c.err').read()[-1500:])
PY
}
RANDOM_SUFFIX=1 run C2 --steps 20 --warmup 5 --no-stages --no-frame
RANDOM_SUFFIX=2 run C3 --steps 6 --warmup 3
RANDOM_SUFFIX=3 run C4 --steps 4 --warmup 3
RANDOM_SUFFIX=4 run C5 --steps 6 --warmup 3

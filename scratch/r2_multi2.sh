#!/bin/bash
# usage: r2_multi2.sh N   -> gpurun_out/r2k_n${N}_${config}.json   (C3 and C4 only: the configs whose kernels changed late in round 2)
N=$1
mkdir -p gpurun_out
run() { c=$1; shift
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --config $c "$@" > gpurun_out/r2k_n${N}_$c.json 2> gpurun_out/r2k_n${N}_$c.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2971$SUF bench.py --gpus $N --config $c "$@" > gpurun_out/r2k_n${N}_$c.json 2> gpurun_out/r2k_n${N}_$c.err
  fi
  echo "$c rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2k_n${N}_$c.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","warmup")}, d.get("e2e",{}).get("value"), d.get("value_hot_path"), d.get("prune"))
except Exception as e:
    print("no json", e); print(open('gpurun_out/r2k_n${N}_$c.err').read()[-1500:])
PY
}
SUF=2 run C3 --steps 6 --warmup 3
SUF=3 run C4 --steps 4 --warmup 3

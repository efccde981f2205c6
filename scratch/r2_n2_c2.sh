#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_n2_C2.json 2> gpurun_out/r2l_n2_C2.err
echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2l_n2_C2.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","warmup","gpu_launches")}, d.get("e2e",{}).get("value"), d.get("value_hot_path"))
except Exception as e:
    print("no json", e); print(open('gpurun_out/r2l_n2_C2.err').read()[-2000:])
PY

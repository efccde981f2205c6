#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/bench_stall_rank*.txt
for rep in 1 2; do
NSVF_BENCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --no-stages --no-frame > gpurun_out/r2c18.json 2> gpurun_out/r2c18.err
grep -E "pre-sized" gpurun_out/r2c18.err; grep -E "rank . step" gpurun_out/r2c18.err | grep -v "+0 cudaMalloc" | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/r2c18.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'])"
done

#!/bin/bash
for m in base base; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scratch/r2_stall.py $m 2>&1 | grep -E "^rank|^cpu.max"
done

#!/bin/bash
timeout 600 python -m pytest tests/test_intersect_gpu.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2 3; do
  for env in LAZY EAGER; do
    echo "== CUDA_MODULE_LOADING=$env rep $rep"
    CUDA_MODULE_LOADING=$env timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scratch/r2_stall.py base 2>&1 | grep -E "^rank" | sed 's/cpu time of main thread+children/cpu/g' | cut -c1-420
  done
done

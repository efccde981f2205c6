#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_stall_tb_rank*.txt
timeout 300 python -m pytest tests/test_intersect_gpu.py -m gpu -x -q 2>&1 | grep -E "Error|assert|passed|failed" | head -12
for rep in 1 2 3 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scratch/r2_stall.py base 2>&1 | grep -E "^rank" | sed 's/cpu time of main thread+children/cpu/g' | cut -c1-300
done
wc -l gpurun_out/r2_stall_tb_rank*.txt

#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/bench_stall_rank*.txt
timeout 300 python -m pytest tests/test_intersect_gpu.py -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
  NSVF_BENCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 --no-stages --no-frame > gpurun_out/r2c17_$rep.json 2> gpurun_out/r2c17_$rep.err
  python - <<PY
import re,json
err=open('gpurun_out/r2c17_$rep.err').read()
pts=re.findall(r"rank (\d) step (\d+) from_host=(\w+): host ([\d.]+) ms \(\+(\d+) cudaMalloc\)", err)
slow=[p for p in pts if float(p[3])>60]
print("rep $rep slow steps:", slow)
d=json.loads(open('gpurun_out/r2c17_$rep.json').read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"])
PY
done

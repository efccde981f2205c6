#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -3
run() { # $1 = config, $2 = steps, $3 = warmup, extra
  c=$1; shift
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2950$RANDOM_SUFFIX bench.py --gpus 2 --config $c "$@" > gpurun_out/r2c12_n2_$c.json 2> gpurun_out/r2c12_n2_$c.err
  echo "$c rc=$?"; tail -c 900 gpurun_out/r2c12_n2_$c.json; echo; grep -v Warning gpurun_out/r2c12_n2_$c.err | tail -4
}
RANDOM_SUFFIX=1 NSVF_BENCH_TRACE=1 run C2 --steps 20 --warmup 5 --no-stages
RANDOM_SUFFIX=2 run C3 --steps 4 --warmup 3
RANDOM_SUFFIX=3 run C4 --steps 4 --warmup 3
RANDOM_SUFFIX=4 run C5 --steps 4 --warmup 3

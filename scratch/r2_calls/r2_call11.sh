#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2c11_bench.err
for c in C3 C4 C5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2c11_$c.json 2> gpurun_out/r2c11_$c.err; echo "$c rc=$?"; tail -c 1200 gpurun_out/r2c11_$c.json; tail -3 gpurun_out/r2c11_$c.err
done

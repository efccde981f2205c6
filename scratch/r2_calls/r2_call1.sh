#!/bin/bash
# round 2, call 1: state of the round-1 code on the new scene sizes + the reference-on-GPU legs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?" >> gpurun_out/r2c1_bench.err
tail -3 gpurun_out/r2c1_pytest.log; tail -c 3000 gpurun_out/r2c1_bench.json; tail -20 gpurun_out/r2c1_bench.err

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2g_pytest.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2g_bench.err

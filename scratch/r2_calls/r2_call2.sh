#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_march_gpu.py tests/test_golden_gpu.py -x -q > gpurun_out/r2c2_march.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_march.log
tail -40 gpurun_out/r2c2_march.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_pytest.log
tail -15 gpurun_out/r2c2_pytest.log

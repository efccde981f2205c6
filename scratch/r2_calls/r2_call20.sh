#!/bin/bash
timeout 600 python -m pytest tests/test_sampling_gpu.py tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scratch/r2_tri.py 2>&1 | grep -v Warn | head -8

#!/bin/bash
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
for v in 1 2 3; do NSVF_TRI_BWD=$v timeout 300 python scratch/r2_tri.py 2>&1 | grep -v Warn | head -8; done

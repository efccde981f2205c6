#!/bin/bash
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
for pm in 1 0; do for v in 1 2; do echo "perm=$pm"; NSVF_TRI_PERM=$pm NSVF_TRI_BWD=$v timeout 300 python scratch/r2_tri.py 2>&1 | grep -E "NSVF_TRI|bwd"; done; done

#!/bin/bash
python -m pytest tests/test_march_gpu.py tests/test_configs_gpu.py tests/test_golden_gpu.py tests/test_ops_gpu.py -m gpu -x -q 2>&1 | tail -2
NSVF_PROFILE_PY=1 python scratch/r2_frame_prof.py trivial 5 2>&1 | head -30

#!/bin/bash
python -m pytest tests/test_sampling_gpu.py tests/test_march_gpu.py tests/test_intersect_gpu.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -6
python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"
NSVF_PROFILE_PY=1 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep -v Warning | grep "function calls\|forward\|synchronize\|window\|fn\b" | head -20

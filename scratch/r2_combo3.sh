#!/bin/bash
python -m pytest tests/test_march_gpu.py tests/test_sampling_gpu.py -m gpu -x -q 2>&1 | tail -6
echo "eager / lazy (stream) frame:"
NSVF_LAZY=0 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"
NSVF_LAZY=1 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"
NSVF_LAZY=1 NSVF_PLANE_BLOCK=32 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"
NSVF_LAZY=1 NSVF_PLANE_BLOCK=128 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"

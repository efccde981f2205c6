#!/bin/bash
# one GPU call: intersection + march tests, walk timings, C3 frame timing with / without launch-ahead and lattice walk
python -m pytest tests/test_intersect_gpu.py tests/test_march_gpu.py -m gpu -x -q 2>&1 | tail -15
python scratch/r2_walk.py 2>&1 | grep -v Warning | tail -8
python scratch/r2_svo.py 2>&1 | grep nodes
echo "frame: ahead+walk / no-ahead+walk / ahead+tree"
NSVF_MARCH_AHEAD=1 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"
NSVF_MARCH_AHEAD=0 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"
NSVF_AABB_NO_GRID=1 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"

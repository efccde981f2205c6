#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scratch/r2_frame_prof.py trivial 5 2>&1 | head -2
timeout 300 python scratch/r2_frame_prof.py mlp 2 2>&1 | head -2

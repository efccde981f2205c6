#!/bin/bash
echo "lazy sampling, plane block 64 / 128 / 192 / 256:"
for b in 64 128 192 256; do NSVF_LAZY=1 NSVF_PLANE_BLOCK=$b python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"; done
echo "eager, plane block 32 / 128:"
for b in 32 128; do NSVF_PLANE_BLOCK=$b python scratch/r2_frame_prof.py trivial 5 2>&1 | grep "^frame"; done
NSVF_PROFILE_PY=1 python scratch/r2_frame_prof.py trivial 5 2>&1 | grep -v Warning | head -70

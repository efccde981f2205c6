#!/bin/bash
for cfg in "1 64" "1 128" "1 256" "0 64" "0 128"; do set -- $cfg; echo "lazy=$1 block=$2"; NSVF_LAZY=$1 NSVF_PLANE_BLOCK=$2 timeout 300 python scratch/r2_frame_prof.py trivial 5 2>&1 | head -1; done

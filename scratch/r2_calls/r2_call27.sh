#!/bin/bash
for v in 2 1 4 5 6; do for s in 1 3; do NSVF_TRI_VARIANT=$v NSVF_TRI_SNAP=$s python scratch/r2_tri2.py 2>&1 | grep VARIANT; done; done
for bps in 12 16 32; do NSVF_TRI_BPS=$bps python scratch/r2_tri2.py 2>&1 | grep VARIANT; done
for b in 1 2 3; do for pm in 0 1; do for s in 0 1; do NSVF_TRI_BWD=$b NSVF_TRI_PERM=$pm NSVF_TRI_SNAP=$s python scratch/r2_tri2.py 2>&1 | grep VARIANT; done; done; done

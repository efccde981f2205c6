#!/bin/bash
for n in 4096 2048 512 64 0; do NSVF_AABB_SMEM_NODES=$n python scratch/r2_aabb.py 2>&1 | grep SMEM; done
for c in 128 64; do NSVF_AABB_LIST_CAP=$c python scratch/r2_aabb.py 2>&1 | grep SMEM; done

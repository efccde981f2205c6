#!/bin/bash
timeout 300 python scratch/r2_step_trace.py trivial 2>&1 | tail -3
echo ---- expandable
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 300 python scratch/r2_step_trace.py trivial 2>&1 | tail -3

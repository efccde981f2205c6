#!/bin/bash
timeout 900 python -m pytest tests/test_march_gpu.py tests/test_sampling_gpu.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -12
timeout 300 python scratch/r2_frame_prof.py trivial 5 2>&1 | head -2
timeout 300 python scratch/r2_frame_prof.py mlp 2 2>&1 | head -2

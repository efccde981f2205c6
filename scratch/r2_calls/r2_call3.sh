#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_configs_gpu.py tests/test_march_gpu.py -x -q 2>&1 | tail -5
NSVF_PROFILE_PY=1 timeout 300 python scratch/r2_frame_prof.py trivial 5 > gpurun_out/r2c3_frame.txt 2>&1
head -60 gpurun_out/r2c3_frame.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c3_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2c3_ncu.log 2>&1
tail -2 gpurun_out/r2c3_ncu.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_march_gpu.py tests/test_configs_gpu.py tests/test_golden_gpu.py -x -q 2>&1 | tail -4
timeout 300 python scratch/r2_frame_prof.py trivial 5 2>&1 | head -3
timeout 300 python scratch/r2_frame_prof.py mlp 2 2>&1 | head -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c5_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2c5_ncu.log 2>&1
python scratch/launch_summary.py gpurun_out/r2c5_frame_launches.csv 3 16

#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c24_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2c24_ncu.log 2>&1
python scratch/launch_summary.py gpurun_out/r2c24_frame_launches.csv 3 14

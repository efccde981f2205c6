#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2f_ncu_frame.log 2>&1
python scratch/launch_summary.py gpurun_out/r2f_frame_launches.csv 3 24
timeout 300 python scratch/r2_frame_prof.py trivial 5 2>&1 | head -1
timeout 300 python scratch/r2_frame_prof.py mlp 3 2>&1 | head -1

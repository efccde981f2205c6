#!/bin/bash
mkdir -p gpurun_out
NSVF_PROFILE_PY=1 timeout 300 python scratch/r2_step_prof.py trivial 20 > gpurun_out/r2c7_step.txt 2>&1
head -70 gpurun_out/r2c7_step.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c7_step_launches.csv python scratch/r2_step_prof.py trivial 1 > gpurun_out/r2c7_ncu.log 2>&1
python scratch/launch_summary.py gpurun_out/r2c7_step_launches.csv 4 40

#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_c3_frame.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2d_launches_c3_frame.log 2>&1
tail -2 gpurun_out/r2d_launches_c3_frame.log

#!/bin/bash
mkdir -p gpurun_out
# full ncu capture of the big one-shot kernels of the C3 frame (second frame: -s skips the warm frame's launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"inverse_cdf_warp_kernel|march_transpose_kernel|aabb_intersect_kernel|march_composite_fwd_kernel" -s 6 -c 6 -o gpurun_out/r2c9_oneshot -f python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2c9_a.log 2>&1
tail -2 gpurun_out/r2c9_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"march_compact_kernel|march_epilogue_kernel|trilinear_fwd" -s 300 -c 6 -o gpurun_out/r2c9_window -f python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2c9_b.log 2>&1
tail -2 gpurun_out/r2c9_b.log
ls -la gpurun_out/*.ncu-rep

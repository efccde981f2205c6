#!/bin/bash
# round 2 evidence run (final code): full GPU test suite, the driver's bench commands, launch lists, ncu --set full of the new kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2h_smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2h_pytest.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2h_ncu_frame.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_step_launches.csv python scratch/r2_step_prof.py trivial 1 > gpurun_out/r2h_ncu_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-frame --no-stages --no-cpu-baseline --no-ref-gpu > gpurun_out/r2h_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"grid_walk_kernel|inverse_cdf_plan_kernel|inverse_cdf_stream_kernel|march_composite_fwd_kernel" -s 36 -c 18 -o gpurun_out/r2h_oneshot -f python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2h_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"march_compact_kernel|march_epilogue_kernel|trilinear_fwd" -s 560 -c 6 -o gpurun_out/r2h_window -f python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2h_b.log 2>&1
ls -la gpurun_out/r2h_*

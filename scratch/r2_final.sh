#!/bin/bash
# round 2 evidence run: full GPU test suite, the driver's bench command, launch lists of the C3 frame and of the C2 step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2f_smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2f_pytest.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2f_ncu_frame.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_step_launches.csv python scratch/r2_step_prof.py trivial 1 > gpurun_out/r2f_ncu_step.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-frame --no-stages --no-cpu-baseline --no-ref-gpu > gpurun_out/r2f_ncu_bench.log 2>&1
ls -la gpurun_out/r2f_*

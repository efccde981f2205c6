#!/bin/bash
# round 2, final code: full GPU test suite, the driver's two bench commands, launch list of the C3 hot-path frame
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2j_smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2j_pytest.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j_frame_launches.csv python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2j_ncu_frame.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"march_count_kernel|march_compact_kernel|march_epilogue_kernel" -s 552 -c 6 -o gpurun_out/r2j_window -f python scratch/r2_frame_prof.py trivial 1 > gpurun_out/r2j_c.log 2>&1
ls -la gpurun_out/r2j_*

mkdir -p gpurun_out
python tests/perf/cublas_emulation_check.py > gpurun_out/r1b_cublas_emulation.txt 2>gpurun_out/err.txt; python tests/perf/cublas_emulation_check.py simt >> gpurun_out/r1b_cublas_emulation.txt 2>>gpurun_out/err.txt; cat gpurun_out/r1b_cublas_emulation.txt; tail -5 gpurun_out/err.txt
python -m pytest tests/test_field_gpu.py -q -x 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-frame --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1b_bench_emu.json; cat gpurun_out/r1b_bench_emu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r1b_launches_step2.csv python bench.py --steps 1 --warmup 3 --no-frame --no-stages --no-cpu-baseline > gpurun_out/r1b_launches_step2.log 2>&1

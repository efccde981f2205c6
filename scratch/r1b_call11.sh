mkdir -p gpurun_out
python -m pytest tests/test_field_gpu.py -q -x 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 --no-frame --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1b_bench3.json; cut -c1-400 gpurun_out/r1b_bench3.json
ncu --set full --clock-control none --import-source on -k regex:ln_relu_bwd -c 2 -o gpurun_out/r1b_ln_bwd python tests/perf/ln_ops.py > /dev/null 2>&1

mkdir -p gpurun_out
python -m pytest tests/test_field_gpu.py tests/test_ops_gpu.py -q -x 2>&1 | tail -15
for run in 6 9; do TRI_ONLY=1 python tests/perf/time_ops.py 40000000 $run 2>&1 | tail -2; done
NSVF_TRI_SNAP=0 TRI_ONLY=1 python tests/perf/time_ops.py 40000000 6 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-frame --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1b_bench_ln.json; cat gpurun_out/r1b_bench_ln.json

mkdir -p gpurun_out
python tests/perf/ln_ops.py
ncu --set full --clock-control none --import-source on -k regex:ln_ -c 8 -o gpurun_out/r1b_ln python tests/perf/ln_ops.py > /dev/null 2>&1
python bench.py --steps 5 --warmup 3 --no-frame --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r1b_bench2.json; cut -c1-2500 gpurun_out/r1b_bench2.json

mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9500 -c 6500 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-frame --no-stages --no-cpu-baseline > gpurun_out/r1b_launches_bench.log 2>&1
tail -c 300 gpurun_out/r1b_launches_bench.log; wc -l gpurun_out/r1b_launches_bench.csv

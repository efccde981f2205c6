set -x
mkdir -p gpurun_out
python -m pytest tests -q -x -m gpu -k "trilinear or embed or interp" 2>&1 | tail -3
for snap in 0 1; do for run in 6 9; do TRI_ONLY=1 NSVF_TRI_SNAP=$snap python tests/perf/time_ops.py 40000000 $run 2>&1 | tail -3; done; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1b_launches_step.csv python bench.py --steps 1 --warmup 3 --no-frame --no-stages --no-cpu-baseline > gpurun_out/r1b_launches_step.log 2>&1
tail -2 gpurun_out/r1b_launches_step.log

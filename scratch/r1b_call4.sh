mkdir -p gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_golden_gpu.py tests/test_configs_gpu.py -q -x 2>&1 | tail -4
for run in 6 9; do TRI_ONLY=1 python tests/perf/time_ops.py 40000000 $run 2>&1 | tail -1; done
for bps in 3 4 5; do NSVF_TRI_BWD_BPS=$bps TRI_ONLY=1 python tests/perf/time_ops.py 40000000 6 2>&1 | tail -1; done
NSVF_TRI_BWD=2 TRI_ONLY=1 python tests/perf/time_ops.py 40000000 6 2>&1 | tail -1
TRI_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:trilinear_bwd -c 2 -o gpurun_out/r1b_tri_bwd_v3 python tests/perf/time_ops.py 40000000 6 > /dev/null 2>&1

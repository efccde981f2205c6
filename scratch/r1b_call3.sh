mkdir -p gpurun_out
python -m pytest tests/test_field_gpu.py tests/test_ops_gpu.py tests/test_golden_gpu.py -q -x 2>&1 | tail -4
for snap in 0 1 2 3; do NSVF_TRI_SNAP=$snap TRI_ONLY=1 python tests/perf/time_ops.py 40000000 6 2>&1 | tail -3; done
TRI_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:trilinear_bwd -c 2 -o gpurun_out/r1b_tri_bwd python tests/perf/time_ops.py 40000000 6 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep

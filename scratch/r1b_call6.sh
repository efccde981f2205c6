for gen in 2 6; do NSVF_TRI_BWD=$gen TRI_ONLY=1 python tests/perf/time_ops.py 40000000 6 2>&1 | tail -1; done

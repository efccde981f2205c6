#!/bin/bash
timeout 600 python -m pytest tests/test_intersect_gpu.py tests/test_golden_gpu.py -m gpu -x -q 2>&1 | tail -2
python tests/perf/ref_clib_compare.py 2>&1 | grep -E "^C2" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[3:]); print({k:d[k] for k in d if 'aabb' in k or 'svo' in k})"

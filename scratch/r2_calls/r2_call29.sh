#!/bin/bash
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests/test_sampling_gpu.py tests/test_march_gpu.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -2
python scratch/r2_frame_prof.py trivial 5 2>&1 | head -1
python tests/perf/ref_clib_compare.py 2>&1 | grep -E "^C[23]" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[3:]); print(l[:2], {k:d[k] for k in d if 'inverse' in k})"

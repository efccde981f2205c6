#!/bin/bash
for m in walk tree walk tree; do
  if [ $m = tree ]; then export NSVF_AABB_NO_GRID=1; else unset NSVF_AABB_NO_GRID; fi
  python bench.py --gpus 1 --steps 20 --warmup 5 --no-frame --no-cpu-baseline --no-ref-gpu --no-stages 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', l['ms_per_step'], l['e2e']['ms_per_step'], l['host_enqueue_ms_per_step'], l['gpu_launches'], l['roofline_hot_path']['kernel_ms'])"
done

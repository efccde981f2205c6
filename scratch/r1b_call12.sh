python -m pytest tests/test_field_gpu.py -q -x -k graphed 2>&1 | tail -15
python bench.py --steps 8 --warmup 3 --no-frame --no-cpu-baseline --no-stages 2>&1 | tail -3 | cut -c1-1500

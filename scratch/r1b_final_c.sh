python -m pytest tests/test_field_gpu.py tests/test_configs_gpu.py -x -q 2>&1 | tail -2
python bench.py --steps 6 --warmup 3 --no-frame --no-stages 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['cuda_graph_replays'], d['roofline']['frac'], d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300

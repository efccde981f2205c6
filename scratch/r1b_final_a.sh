mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err; tail -c 6000 gpurun_out/r1b_bench_n1.json; tail -3 gpurun_out/r1b_bench_n1.err

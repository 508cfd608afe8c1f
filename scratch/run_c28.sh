python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_c28.json 2> gpurun_out/bench_c28.err; tail -2 gpurun_out/bench_c28.err

python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; tail -2 gpurun_out/bench_last.err

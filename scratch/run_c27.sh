python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c27_$i.json 2> gpurun_out/bench_c27_$i.err; tail -2 gpurun_out/bench_c27_$i.err; done

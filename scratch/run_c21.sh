python -m pytest tests/test_gpu_ops.py tests/test_gpu_cogmen.py tests/test_gpu_dgcn.py -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c21.json 2> gpurun_out/bench_c21.err; tail -2 gpurun_out/bench_c21.err

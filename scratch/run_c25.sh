python -m pytest tests/test_gpu_ops.py tests/test_gpu_cogmen.py -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c25_S5.json 2> gpurun_out/bench_c25_S5.err; tail -2 gpurun_out/bench_c25_S5.err
ERCG_LIB_PATH=$PWD/scratch/variants/S4.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c25_S4.json 2> gpurun_out/bench_c25_S4.err; tail -2 gpurun_out/bench_c25_S4.err

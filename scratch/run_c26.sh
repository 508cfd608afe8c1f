for r in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c26_new$r.json 2> gpurun_out/bench_c26_new$r.err
ERCG_LIB_PATH=$PWD/scratch/variants/OLD.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c26_old$r.json 2> gpurun_out/bench_c26_old$r.err
done

python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scratch/bench_gemm.py > gpurun_out/gemm_c5.txt 2>&1; cat gpurun_out/gemm_c5.txt
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_c5_$i.json 2> gpurun_out/bench_c5_$i.err; tail -2 gpurun_out/bench_c5_$i.err; done

python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for f in base nnR6Q2 tnR8 tnR6Q4; do echo "== $f"; ERCG_LIB_PATH=$PWD/scratch/variants/$f.so python scratch/bench_gemm.py 2>&1 | grep -E "K=1443|K1=1443|K= 100 N= 400|K1= 100 N1= 400|K= 400"; done > gpurun_out/variants_c6.txt 2>&1
cat gpurun_out/variants_c6.txt
for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_c6_$i.json 2> gpurun_out/bench_c6_$i.err; tail -2 gpurun_out/bench_c6_$i.err; done

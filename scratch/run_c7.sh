cat /sys/fs/cgroup/cpu.max 2>/dev/null; grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null
for f in base R3Q3 R4Q2 R3Q2 R2Q2; do echo "== $f"; ERCG_LIB_PATH=$PWD/scratch/variants/$f.so python scratch/bench_gemm.py 2>&1 | grep -E "K=1443|K1=1443|K= 100 N= 400|K1= 100 N1= 400|K= 400"; done > gpurun_out/variants_c7.txt 2>&1
cat gpurun_out/variants_c7.txt
grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null
for i in 1 2 3; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_c7_$i.json 2> gpurun_out/bench_c7_$i.err; tail -2 gpurun_out/bench_c7_$i.err; grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null; done

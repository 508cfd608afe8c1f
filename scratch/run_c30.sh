python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c30_base.json 2>/dev/null
for f in CT_R4B3 CT_R2B3 CT_R2B4 TN_G4 TN_G16; do ERCG_LIB_PATH=$PWD/scratch/variants/$f.so python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c30_$f.json 2>/dev/null; done

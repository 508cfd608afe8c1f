python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py -x -q 2>&1 | tail -3
for f in G4D4 G4D1 G8D1 G8D2 G16D2; do echo "== $f"; ERCG_LIB_PATH=$PWD/scratch/variants/$f.so python scratch/bench_gemm.py 2>&1 | grep -E "K1="; done > gpurun_out/variants_c16.txt 2>&1
cat gpurun_out/variants_c16.txt

python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_ops.py tests/test_gpu_cogmen.py -x -q 2>&1 | tail -3
python scratch/bench_gemm.py 2>&1 | head -8

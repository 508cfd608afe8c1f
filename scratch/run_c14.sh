python -m pytest tests/test_gpu_gemm_tc.py -x -q 2>&1 | tail -15
python scratch/bench_gemm.py 2>&1 | head -8

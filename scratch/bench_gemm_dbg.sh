for d in 0 32 64 65 79; do echo "== ERCG_TC_DBG=$d"; ERCG_TC_DBG=$d python scratch/bench_gemm.py 2>&1 | grep -E "^  K=" ; done

for rep in 1 2; do for f in scratch/variants/*.so; do echo "== $f"; ERCG_LIB_PATH=$PWD/$f python scratch/bench_gemm.py 2>&1 | grep -E "K=1443|K1=1443|K1= 100 N1= 900"; done; done

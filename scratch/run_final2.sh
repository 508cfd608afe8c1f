python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
( time python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err ) 2>&1 | grep real

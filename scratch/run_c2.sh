set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python bench.py --profile-step --total-utts 262144 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:_kernel -o gpurun_out/prof_c2 -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_c2.log 2>&1
tail -3 gpurun_out/ncu_c2.log

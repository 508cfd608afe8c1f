set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -3 gpurun_out/bench_c3.err
python bench.py --profile-step --total-utts 262144 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -k regex:'attn_.*tile|gemm_tc_nn|gemm_tc_tn|graphify|gather' -o gpurun_out/prof_c3a -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_c3a.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tc_nn_kernel -s 4 -c 1 -o gpurun_out/prof_c3b -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_c3b.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attn_fwd_tile -c 1 -o gpurun_out/prof_c3c -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_c3c.log 2>&1
ls -la gpurun_out

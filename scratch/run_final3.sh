python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; tail -2 gpurun_out/bench_r01e.err
python bench.py --profile-step > gpurun_out/plain_prof.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -k regex:'gather_fwd_flat' -c 1 -o gpurun_out/prof_gf -f python bench.py --profile-step > gpurun_out/ncu_gf.log 2>&1
python profiles/ncu_summary.py full gpurun_out/prof_gf.ncu-rep gpurun_out/r01e_ncu_full_gather_fwd_flat_1M.csv; rm -f gpurun_out/prof_gf.ncu-rep

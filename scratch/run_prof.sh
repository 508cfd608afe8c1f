set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -2 gpurun_out/bench_r01c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01c_reference.json 2> gpurun_out/bench_r01c_reference.err; tail -2 gpurun_out/bench_r01c_reference.err
# launch list of one warm step (every kernel, ours and ATen's)
python bench.py --profile-step --total-utts 262144 > gpurun_out/plain_prof.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01c_raw.csv \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_list.log 2>&1
python profiles/ncu_summary.py list gpurun_out/launches_r01c_raw.csv gpurun_out/r01c_launches_262k.csv
# full sets: tensor-core GEMMs, graph kernels
ncu --set full --clock-control none --profile-from-start off -k regex:'gemm_tc' -o gpurun_out/prof_gemm -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_gemm.log 2>&1
python profiles/ncu_summary.py full gpurun_out/prof_gemm.ncu-rep gpurun_out/r01c_ncu_full_gemm_262k.csv; rm -f gpurun_out/prof_gemm.ncu-rep
ncu --set full --clock-control none --profile-from-start off -k regex:'attn_.*tile|graphify|rel_|gather|skinny|col_partials|colsum|bn_|mask_pos|ce_' -o gpurun_out/prof_graph -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_graph.log 2>&1
python profiles/ncu_summary.py full gpurun_out/prof_graph.ncu-rep gpurun_out/r01c_ncu_full_graph_262k.csv; rm -f gpurun_out/prof_graph.ncu-rep
# dominant kernel at FULL size (2^20 utterances): DRAM traffic per launch
ncu --set full --clock-control none --profile-from-start off -k regex:'gemm_tc_tn_kernel' -s 3 -c 1 -o gpurun_out/prof_dom -f \
  python bench.py --profile-step > gpurun_out/ncu_dom.log 2>&1
python profiles/ncu_summary.py full gpurun_out/prof_dom.ncu-rep gpurun_out/r01c_ncu_full_dominant_1M.csv; rm -f gpurun_out/prof_dom.ncu-rep
rm -f gpurun_out/launches_r01c_raw.csv
ls -la gpurun_out

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -2 gpurun_out/bench_r01d.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01d_reference.json 2> gpurun_out/bench_r01d_reference.err
# launch list of one warm step of the DEFAULT workload (2^20 utterances): every kernel, ours and ATen's
python bench.py --profile-step > gpurun_out/plain_prof.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_raw.csv \
  python bench.py --profile-step > gpurun_out/ncu_list.log 2>&1
python profiles/ncu_summary.py list gpurun_out/launches_raw.csv gpurun_out/r01d_launches_1M.csv; rm -f gpurun_out/launches_raw.csv
# ncu --set full: tensor-core GEMMs (full size), then the graph / elementwise kernels (full size)
ncu --set full --clock-control none --profile-from-start off -k regex:'gemm_tc' -o gpurun_out/prof_gemm -f \
  python bench.py --profile-step > gpurun_out/ncu_gemm.log 2>&1
python profiles/ncu_summary.py full gpurun_out/prof_gemm.ncu-rep gpurun_out/r01d_ncu_full_gemm_1M.csv; rm -f gpurun_out/prof_gemm.ncu-rep
ncu --set full --clock-control none --profile-from-start off -k regex:'attn_.*tile|graphify|rel_|gather|skinny|col_stream|cls_tail|bn_|ce_' -o gpurun_out/prof_graph -f \
  python bench.py --profile-step > gpurun_out/ncu_graph.log 2>&1
python profiles/ncu_summary.py full gpurun_out/prof_graph.ncu-rep gpurun_out/r01d_ncu_full_graph_1M.csv; rm -f gpurun_out/prof_graph.ncu-rep
ls -la gpurun_out

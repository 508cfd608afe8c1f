python bench.py --profile-step --total-utts 262144 > gpurun_out/plain_c24.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'attn_fwd_tile|attn_bwd_src_tile|attn_bwd_dst_tile' -o gpurun_out/prof_c24 -f \
  python bench.py --profile-step --total-utts 262144 > gpurun_out/ncu_c24.log 2>&1
ls -la gpurun_out/prof_c24.ncu-rep

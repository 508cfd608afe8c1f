python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/bench_r01e_n2.json 2> gpurun_out/bench_r01e_n2.err; tail -2 gpurun_out/bench_r01e_n2.err

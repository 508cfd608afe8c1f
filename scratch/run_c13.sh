python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_c13_1.json 2> gpurun_out/bench_c13_1.err; tail -2 gpurun_out/bench_c13_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 2 > gpurun_out/bench_c13_n2.json 2> gpurun_out/bench_c13_n2.err; tail -3 gpurun_out/bench_c13_n2.err

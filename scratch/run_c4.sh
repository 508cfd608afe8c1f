for i in 1 2; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c4_$i.json 2> gpurun_out/bench_c4_$i.err; done
ERCG_NO_CENSUS=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_c4_nc.json 2> gpurun_out/bench_c4_nc.err
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv; nproc; cat /proc/loadavg

for cs in 1 2 3; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 4 --copy-streams $cs > gpurun_out/bench_c19_$cs.json 2> gpurun_out/bench_c19_$cs.err; tail -2 gpurun_out/bench_c19_$cs.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c19_$cs.json')); print($cs, d['e2e'])"; done
python -m pytest tests/test_gpu_ops.py -q -k feeder 2>&1 | tail -2

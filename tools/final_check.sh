#!/bin/bash
# What the driver runs at round end, on one B200: GPU tests, smoke, both bench arms; optional ncu launch list (LIST=1).
tag=${1:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${tag}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${tag}_smoke.log
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${tag}_bench_n1.json
if [ -n "$REF" ]; then
  python bench.py --impl reference > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_bench_reference_arm.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_reference_arm.json
fi
if [ -n "$LIST" ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/${tag}_launches_raw.csv python bench.py --profile-step > gpurun_out/${tag}_ncu_list.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_ncu_list.log
fi

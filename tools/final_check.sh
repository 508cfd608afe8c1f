python -m pytest tests -m gpu -x -q > gpurun_out/r02zi_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02zi_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02zi_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02zi_smoke.log
python bench.py > gpurun_out/r02zi_bench_n1.json 2> gpurun_out/r02zi_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02zi_bench_n1.json
python bench.py --impl reference > gpurun_out/r02zi_bench_reference_arm.json 2> gpurun_out/r02zi_bench_reference_arm.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02zi_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02zi_launches_raw.csv python bench.py --profile-step > gpurun_out/r02zi_ncu_list.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02zi_ncu_list.log

set -x
timeout 300 python benchmarks/step_profile.py 3 > gpurun_out/s2_step_profile_final.txt 2>&1
timeout 600 python bench.py > gpurun_out/s2_bench_final.json 2> gpurun_out/s2_bench_final.err; echo bench rc=$?
timeout 300 python benchmarks/kernel_bench.py simam > gpurun_out/s2_kb_simam.log 2>&1; echo kb rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_bench_launches_final.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline > gpurun_out/s2_ncu_bench.log 2>&1; echo ncu rc=$?
tail -c 400 gpurun_out/s2_bench_final.json; wc -l gpurun_out/s2_bench_launches_final.csv

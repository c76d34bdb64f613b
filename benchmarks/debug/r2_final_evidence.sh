#!/bin/bash
# Round-2 final evidence (one gpurun call): tests, bench lines, microbenchmarks, ncu launch list + full captures.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
set -x
(cd tests && timeout 900 python -m pytest . -x -q -m gpu 2>&1 | tail -4) > gpurun_out/r2f_pytest_gpu.log 2>&1; cat gpurun_out/r2f_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo ref rc=$?
timeout 200 python benchmarks/kernel_bench.py attn --engine tcgen05 > gpurun_out/r2f_kernel_bench_attn.jsonl 2>&1
timeout 200 python benchmarks/kernel_bench.py attn_long > gpurun_out/r2f_kernel_bench_attn_long.jsonl 2>&1
timeout 300 python benchmarks/kernel_bench.py simam --no-workspace > gpurun_out/r2f_kernel_bench_simam.jsonl 2>&1
timeout 300 python benchmarks/linear_bench.py > gpurun_out/r2f_linear_bench.jsonl 2>&1
timeout 300 python benchmarks/config_bench.py 2 > gpurun_out/r2f_config2.jsonl 2> gpurun_out/r2f_config2.err
timeout 300 python benchmarks/config_bench.py 5 > gpurun_out/r2f_config5.jsonl 2> gpurun_out/r2f_config5.err
timeout 300 python benchmarks/step_profile.py 3 > gpurun_out/r2f_step_profile.txt 2>&1
# ncu: launch list of one eager train step (the timed region replays this step as a CUDA graph), then full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2f_ncu_bench.log 2>&1; echo ncu rc=$?
for t in "attn s3" "attn s1" "linear s3" "attn_long x"; do
  n=$(echo $t | tr ' ' '_')
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2f_$n python benchmarks/ncu_targets.py $t > gpurun_out/r2f_ncu_$n.log 2>&1; echo "ncu $t rc=$?"
done
ls -la gpurun_out/r2f_*

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_linear_gpu.py -q -m gpu -x > gpurun_out/r2_pytest_linear.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_linear.log
tail -25 gpurun_out/r2_pytest_linear.log | cut -c1-220
timeout 600 python benchmarks/linear_bench.py --only fc1 > gpurun_out/r2_linear_bench.jsonl 2> gpurun_out/r2_linear_bench.err
echo "linear_bench rc=$?"; grep dgelu gpurun_out/r2_linear_bench.jsonl; tail -5 gpurun_out/r2_linear_bench.err

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py -q -m gpu -k "wgrad" > gpurun_out/r2_pytest_wgrad.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_wgrad.log
tail -25 gpurun_out/r2_pytest_wgrad.log | cut -c1-250
timeout 600 python benchmarks/linear_bench.py wgrad > gpurun_out/r2_wgrad_bench.jsonl 2> gpurun_out/r2_wgrad_bench.err
echo "wgrad_bench rc=$?"; cat gpurun_out/r2_wgrad_bench.jsonl; tail -5 gpurun_out/r2_wgrad_bench.err

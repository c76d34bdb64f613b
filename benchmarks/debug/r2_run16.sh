#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stripe_attn_gpu.py tests/test_attn_dropout_gpu.py -q -m gpu -x > gpurun_out/r2_pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_attn.log
tail -12 gpurun_out/r2_pytest_attn.log | cut -c1-250
timeout 600 python benchmarks/kernel_bench.py attn --engine tcgen05 > gpurun_out/r2_kernel_bench_attn.jsonl 2> gpurun_out/r2_kernel_bench_attn.err
echo "kb rc=$?"; cat gpurun_out/r2_kernel_bench_attn.jsonl | cut -c1-400; tail -3 gpurun_out/r2_kernel_bench_attn.err

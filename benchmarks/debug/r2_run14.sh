#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_dropout_gpu.py -q -m gpu > gpurun_out/r2_pytest_drop.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_drop.log
tail -40 gpurun_out/r2_pytest_drop.log | cut -c1-250
timeout 900 python -m pytest tests/test_stripe_attn_gpu.py tests/test_models_gpu.py -q -m gpu > gpurun_out/r2_pytest_attn.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_attn.log
tail -5 gpurun_out/r2_pytest_attn.log | cut -c1-250

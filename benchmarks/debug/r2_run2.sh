#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python benchmarks/step_profile.py 2 ops > gpurun_out/r2_step_ops.txt 2>&1
echo rc=$?
sed -n '/--- aten ops/,$p' gpurun_out/r2_step_ops.txt | cut -c1-220 | head -60

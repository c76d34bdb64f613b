#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python benchmarks/ncu_targets.py linear s3 > gpurun_out/r2_ncu_plain.log 2>&1 && python benchmarks/ncu_targets.py linear s1 >> gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_linear_s3 python benchmarks/ncu_targets.py linear s3 > gpurun_out/r2_ncu_lin3.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_linear_s1 python benchmarks/ncu_targets.py linear s1 > gpurun_out/r2_ncu_lin1.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_ncu_lin1.log; ls -la gpurun_out/*.ncu-rep

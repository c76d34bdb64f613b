#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python benchmarks/ncu_targets.py attn s3 > gpurun_out/r2_ncu_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_attn_s3 python benchmarks/ncu_targets.py attn s3 > gpurun_out/r2_ncu_attn.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_attn.log; ls -la gpurun_out/*.ncu-rep

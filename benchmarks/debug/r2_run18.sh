#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python benchmarks/debug/glue_ops.py all > gpurun_out/r2h_glue_ops.txt 2>&1
tail -5 gpurun_out/r2h_glue_ops.txt

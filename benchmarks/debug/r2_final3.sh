#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 85 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo bench rc=$?
tail -c 300 gpurun_out/r2n_bench.err; tail -c 1200 gpurun_out/r2n_bench.json | head -c 600

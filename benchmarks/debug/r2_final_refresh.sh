#!/bin/bash
# Refresh of the round-2 evidence after the last kernel changes (pair mode for 64-token stripes, max chains).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
set -x
(cd tests && timeout 900 python -m pytest . -x -q -m gpu 2>&1 | tail -4) > gpurun_out/r2g_pytest_gpu.log 2>&1; cat gpurun_out/r2g_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2g_smoke.log 2>&1; tail -2 gpurun_out/r2g_smoke.log
timeout 600 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo bench rc=$?
timeout 200 python benchmarks/kernel_bench.py attn --engine tcgen05 > gpurun_out/r2g_kernel_bench_attn.jsonl 2>&1
timeout 200 python benchmarks/kernel_bench.py attn_long > gpurun_out/r2g_kernel_bench_attn_long.jsonl 2>&1
timeout 300 python benchmarks/config_bench.py 5 > gpurun_out/r2g_config5.jsonl 2> gpurun_out/r2g_config5.err
tail -c 600 gpurun_out/r2g_bench.json; cat gpurun_out/r2g_config5.jsonl

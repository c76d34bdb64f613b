#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 300 $TR benchmarks/debug/dp_profile.py > gpurun_out/r2h_dp_profile_fp32.txt 2>&1
timeout 300 $TR benchmarks/debug/dp_profile.py bf16 > gpurun_out/r2h_dp_profile_bf16.txt 2>&1
(cd tests && timeout 600 python -m pytest test_dp_nccl_gpu.py -x -q -m gpu 2>&1 | tail -5) > gpurun_out/r2h_dp_test.log 2>&1
grep -v Warning gpurun_out/r2h_dp_profile_fp32.txt | tail -30; grep "ms/step\|nccl kernels" gpurun_out/r2h_dp_profile_bf16.txt; cat gpurun_out/r2h_dp_test.log

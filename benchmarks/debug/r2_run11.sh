#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -q -m gpu > gpurun_out/r2_pytest_dp.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_dp.log
tail -30 gpurun_out/r2_pytest_dp.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['replicas_in_sync'], d['config'].get('dp_allreduce'))
PY
tail -5 gpurun_out/r2_bench_n2.err

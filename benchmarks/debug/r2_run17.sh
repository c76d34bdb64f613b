#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "bench n8 rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2b.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['replicas_in_sync'], d.get('config4_global_batch_256'))
PY
tail -3 gpurun_out/r2_bench_n2b.err | cut -c1-300

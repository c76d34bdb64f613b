#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layernorm_gpu.py -q -m gpu > gpurun_out/r2_pytest_ln.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_ln.log
tail -30 gpurun_out/r2_pytest_ln.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench6.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'])
print({k:(v['frac'],v['ms_total']) for k,v in d['roofline_all'].items()})
PY
tail -3 gpurun_out/r2_bench6.err

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
tail -12 gpurun_out/r2_pytest.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 --no-gpu-baseline --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel'], d['roofline']['frac'])
for r in d['roofline_shapes']: print(r)
PY
tail -3 gpurun_out/r2_bench4.err

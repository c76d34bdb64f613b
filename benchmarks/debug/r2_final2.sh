#!/bin/bash
# Final round-2 evidence after the deferred final sums / zero arena (one gpurun call, ~5 GPU-minutes).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(cd tests && timeout 400 python -m pytest . -x -q -m gpu 2>&1 | tail -6) > gpurun_out/r2m_pytest_gpu.log 2>&1; cat gpurun_out/r2m_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
timeout 400 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo bench rc=$?
tail -c 400 gpurun_out/r2m_bench.err
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cuda-graph --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2m_ncu_bench.log 2>&1; echo ncu rc=$?
timeout 100 python benchmarks/step_profile.py 3 > gpurun_out/r2m_step_profile.txt 2>&1; echo prof rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"])
PY

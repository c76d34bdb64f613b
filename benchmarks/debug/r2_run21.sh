#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(cd tests && timeout 240 python -m pytest test_layernorm_gpu.py test_models_gpu.py test_linear_gpu.py test_gelu_conv_gpu.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2j_tests.log 2>&1; cat gpurun_out/r2j_tests.log
CSB200_DEFER_SUMS=0 timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2j_bench_immediate.json 2> gpurun_out/r2j_bench_immediate.err; echo rc=$?
timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2j_bench_deferred.json 2> gpurun_out/r2j_bench_deferred.err; echo rc=$?
python - <<'PY'
import json
for n in ("immediate","deferred"):
    try:
        d=json.loads(open(f"gpurun_out/r2j_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"])
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r2j_bench_{n}.err").read()[-1500:])
PY

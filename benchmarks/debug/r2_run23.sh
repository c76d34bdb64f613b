#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(cd tests && timeout 300 python -m pytest test_layernorm_gpu.py test_models_gpu.py test_linear_gpu.py test_gelu_conv_gpu.py test_optim_gpu.py test_capi.py -q -m gpu 2>&1 | tail -25) > gpurun_out/r2l_tests.log 2>&1; cat gpurun_out/r2l_tests.log
timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2l_bench_deferred.json 2> gpurun_out/r2l_bench_deferred.err; echo rc=$?
python - <<'PY'
import json
for n in ("deferred",):
    try:
        d=json.loads(open(f"gpurun_out/r2l_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["gpu_launches"])
    except Exception as e:
        print(n, "FAILED", e); print(open(f"gpurun_out/r2l_bench_{n}.err").read()[-1500:])
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
(cd tests && timeout 120 python -m pytest test_linear_gpu.py -x -q -m gpu -k wgrad 2>&1 | tail -3) > gpurun_out/r2i_wgrad_test.log 2>&1; cat gpurun_out/r2i_wgrad_test.log
for n in 1 2 3; do
  CSB200_WGRAD_PRODUCERS=$n timeout 100 python benchmarks/linear_bench.py wgrad > gpurun_out/r2i_wgrad_bench_p$n.jsonl 2>&1
done
python - <<'PY'
import json
rows={n:[json.loads(l) for l in open(f"gpurun_out/r2i_wgrad_bench_p{n}.jsonl") if l.startswith("{")] for n in (1,2,3)}
for i,r in enumerate(rows[1]):
    print(r["shape"], "cublas", r["cublas_wgrad_only_us"], "+colsum", r["cublas_plus_colsum_us"], "| p1", r["csb200_us"], "p2", rows[2][i]["csb200_us"], "p3", rows[3][i]["csb200_us"])
PY

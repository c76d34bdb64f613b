#!/bin/bash
# round 2, first GPU call: the new parity tests, smoke, bench with the reference legs
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt
nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_models_gpu.py::test_cswin_512_golden_through_the_tcgen05_engines > gpurun_out/r2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest.log
timeout 600 python -m pytest tests/test_models_gpu.py -q -m gpu -k "512_golden or bf16_within" > gpurun_out/r2_pytest_512.log 2>&1
echo "pytest512 rc=$?" >> gpurun_out/r2_pytest_512.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo "bench rc=$?" >> gpurun_out/r2_bench1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
tail -3 gpurun_out/r2_pytest.log; tail -5 gpurun_out/r2_pytest_512.log; tail -2 gpurun_out/r2_smoke.log; tail -c 600 gpurun_out/r2_bench1.err

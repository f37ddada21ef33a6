set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s60_pytest_gpu.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/s60_smoke.log 2>&1
timeout 400 python bench.py --no-convergence > gpurun_out/s60_bench.json 2> gpurun_out/s60_bench.err

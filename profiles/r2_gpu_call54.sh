set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s54_pytest_gpu.txt 2>&1
timeout 600 python bench.py > gpurun_out/s54_bench_train.json 2> gpurun_out/s54_bench_train.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/s54_smoke.log 2>&1

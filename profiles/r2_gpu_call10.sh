set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu 2>&1 | tail -15 > gpurun_out/s10_tests.log
timeout 900 python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err

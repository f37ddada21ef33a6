set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s21_pytest_gpu.txt 2>&1
timeout 300 python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err
AIR_WB_REF=0 timeout 300 python bench.py > gpurun_out/s21_bench_clean.json 2> gpurun_out/s21_bench_clean.err

set -x
mkdir -p gpurun_out
timeout 500 python bench.py > gpurun_out/s36_bench.json 2> gpurun_out/s36_bench.err

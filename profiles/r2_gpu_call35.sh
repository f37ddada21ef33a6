set -x
mkdir -p gpurun_out
AIR_DEBUG_CAPTURE=1 timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/s35_bench.json 2> gpurun_out/s35_bench.err

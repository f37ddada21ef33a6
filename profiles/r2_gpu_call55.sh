set -x
mkdir -p gpurun_out
AIR_BENCH_CONV_INPROC=1 AIR_BENCH_CONV_ITERS=300 timeout 500 python bench.py --steps 10 --warmup 3 > gpurun_out/s55_bench_inproc.json 2> gpurun_out/s55_bench_inproc.err

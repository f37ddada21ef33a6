set -x
mkdir -p gpurun_out
timeout 300 python -c "import bench, json; print(json.dumps(bench.training_convergence((3,), 3000)))" > gpurun_out/s32_conv_standalone.txt 2>&1
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/s32_bench.json 2> gpurun_out/s32_bench.err

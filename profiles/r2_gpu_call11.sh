set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_gpu_check.py > gpurun_out/s11_dp_check.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/s11_bench_n2.json 2> gpurun_out/s11_bench_n2.err
timeout 900 python -m pytest tests/test_gpu_baseline_sizes.py -q -x 2>&1 | tail -15 > gpurun_out/s11_baseline_sizes.log

set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/s56_bench_n4.json 2> gpurun_out/s56_bench_n4.err

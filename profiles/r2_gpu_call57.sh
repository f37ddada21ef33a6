set -x
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/s57_bench_n8.json 2> gpurun_out/s57_bench_n8.err

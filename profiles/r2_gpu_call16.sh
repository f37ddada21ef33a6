set -x
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_gpu_check.py > gpurun_out/s16_dp_check.log 2>&1
echo "dp_check rc=$?" >> gpurun_out/s16_dp_check.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/s16_bench_n2.json 2> gpurun_out/s16_bench_n2.err
echo "bench rc=$?" >> gpurun_out/s16_bench_n2.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/s16_bench_ref_n2.json 2> gpurun_out/s16_bench_ref_n2.err
echo "ref rc=$?" >> gpurun_out/s16_bench_ref_n2.err

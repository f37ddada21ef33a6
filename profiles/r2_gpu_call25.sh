set -x
mkdir -p gpurun_out
for S in 0 1 2 3; do
  timeout 300 python examples/train_synthetic.py --data oracle --train-images 20000 --val-images 1024 --seed $S --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s25_conv_v1_odata_seed$S.log > /dev/null 2>&1
  AIR_WB_REF_ATOMICS=1 timeout 300 python examples/train_synthetic.py --data oracle --train-images 20000 --val-images 1024 --seed $S --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s25_conv_atomics_odata_seed$S.log > /dev/null 2>&1
done

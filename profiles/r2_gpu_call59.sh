set -x
mkdir -p gpurun_out
for S in 0 1 2 3 4 5 6 7 8 9 10 11; do
  timeout 100 python examples/train_synthetic.py --seed $S --iters 25000 --every 5000 --gemm tf32x3 --log gpurun_out/s59_conv_final_seed$S.log > /dev/null 2>&1
done

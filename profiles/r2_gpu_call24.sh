set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_wb_reference_rounding.py -q > gpurun_out/s24_pytest_ref.txt 2>&1
ITERS=41 EVERY=10 MODE=tf32x3 timeout 300 python tests/diag_teacher_forced.py > gpurun_out/s24_teacher_forced_ref.txt 2>&1
for S in 0 1 2 3; do
  timeout 200 python examples/train_synthetic.py --seed $S --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s24_conv_v1_seed$S.log > /dev/null 2>&1
done

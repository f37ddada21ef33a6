set -x
mkdir -p gpurun_out
ITERS=41 EVERY=10 MODE=tf32x3 timeout 300 python tests/diag_teacher_forced.py > gpurun_out/s20_teacher_forced_ref.txt 2>&1
timeout 400 python examples/train_synthetic.py --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s20_conv_ref_x3.log > /dev/null 2>&1
AIR_WB_REF=0 timeout 400 python examples/train_synthetic.py --iters 6000 --every 1000 --gemm tf32x3 --log gpurun_out/s20_conv_clean_x3.log > /dev/null 2>&1

set -x
mkdir -p gpurun_out
AIR_WB_REF_ATOMICS=4 timeout 300 python -m pytest tests/test_gpu_wb_reference_rounding.py -q -s -k "not training" > gpurun_out/s52_pytest_ref_sep4.txt 2>&1
AIR_WB_REF_ATOMICS=4 ITERS=41 EVERY=10 MODE=tf32x3 timeout 300 python tests/diag_teacher_forced.py 2>&1 | cut -c1-110 > gpurun_out/s52_teacher_forced_sep4.txt
for S in 0 1 2 3 4 5 6 7; do
  AIR_WB_REF_ATOMICS=4 timeout 100 python examples/train_synthetic.py --seed $S --iters 25000 --every 5000 --gemm tf32x3 --log gpurun_out/s52_conv_sep4_seed$S.log > /dev/null 2>&1
done

set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/s47_smoke.log 2>&1
N=$(STEPS=2 python tests/diag_train_steps.py | sed -n 's/.*kernels_per_step \([0-9]*\).*/\1/p')
echo "kernels per step: $N" > gpurun_out/s47_kernels_per_step.txt
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*N)) -c $N --csv --log-file gpurun_out/s47_launches_train.csv python tests/diag_train_steps.py > gpurun_out/s47_ncu_list.log 2>&1
STEPS=3 MODE=tf32x3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*N)) -c $N --csv --log-file gpurun_out/s47_launches_train_x3.csv python tests/diag_train_steps.py > gpurun_out/s47_ncu_list_x3.log 2>&1
timeout 200 python tests/diag_step_profile.py > gpurun_out/s47_step_kernels_warm.txt 2>&1
MODE=tf32x3 timeout 200 python tests/diag_step_profile.py > gpurun_out/s47_step_kernels_warm_x3.txt 2>&1
timeout 500 python bench.py > gpurun_out/s47_bench_train.json 2> gpurun_out/s47_bench_train.err
timeout 300 python bench.py --workload st > gpurun_out/s47_bench_st.json 2> gpurun_out/s47_bench_st.err
timeout 300 python bench.py --workload infer > gpurun_out/s47_bench_infer.json 2> gpurun_out/s47_bench_infer.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s47_bench_reference.json 2> gpurun_out/s47_bench_reference.err
timeout 100 python examples/train_synthetic.py --seed 3 --iters 25000 --every 5000 --gemm tf32 --log gpurun_out/s47_conv_tf32_seed3.log > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s47_pytest_gpu.txt 2>&1

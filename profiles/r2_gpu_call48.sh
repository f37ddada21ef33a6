set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -q -k "adam" > gpurun_out/s48_pytest_adam.txt 2>&1
timeout 500 python bench.py --no-convergence > gpurun_out/s48_bench_train.json 2> gpurun_out/s48_bench_train.err
N=76
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*N)) -c $N --csv --log-file gpurun_out/s48_launches_train.csv python tests/diag_train_steps.py > gpurun_out/s48_ncu_list.log 2>&1

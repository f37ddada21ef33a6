set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s51_pytest_gpu.txt 2>&1
timeout 600 python bench.py > gpurun_out/s51_bench_train.json 2> gpurun_out/s51_bench_train.err
timeout 300 python bench.py --workload st > gpurun_out/s51_bench_st.json 2> gpurun_out/s51_bench_st.err
N=76
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*N)) -c $N --csv --log-file gpurun_out/s51_launches_train.csv python tests/diag_train_steps.py > gpurun_out/s51_ncu_list.log 2>&1
STEPS=3 MODE=tf32x3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*N)) -c $N --csv --log-file gpurun_out/s51_launches_train_x3.csv python tests/diag_train_steps.py > gpurun_out/s51_ncu_list_x3.log 2>&1
AIR_PDL=0 timeout 200 python tests/diag_step_profile.py > gpurun_out/s51_step_kernels_warm.txt 2>&1
AIR_PDL=0 MODE=tf32x3 timeout 200 python tests/diag_step_profile.py > gpurun_out/s51_step_kernels_warm_x3.txt 2>&1
mkdir -p /tmp/ncu
STEPS=2 timeout 200 ncu --set full --clock-control none -k regex:st_wb_bwd_ref -s 1 -c 1 -o /tmp/ncu/ref python tests/diag_train_steps.py > gpurun_out/s51_ncu_ref.log 2>&1
ncu -i /tmp/ncu/ref.ncu-rep --page raw --csv > gpurun_out/s51_st_wb_bwd_ref_raw.csv 2>/dev/null
STEPS=2 timeout 200 ncu --set full --clock-control none -k regex:st_compose_steps -s 1 -c 1 -o /tmp/ncu/compose python tests/diag_train_steps.py > gpurun_out/s51_ncu_compose.log 2>&1
ncu -i /tmp/ncu/compose.ncu-rep --page raw --csv > gpurun_out/s51_st_compose_raw.csv 2>/dev/null

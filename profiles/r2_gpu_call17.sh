set -x
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/s17_smoke.log 2>&1
# launch list of the final schedule: 2 warm-up steps skipped (76 kernels per step), one full step listed
STEPS=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 152 -c 76 --csv --log-file gpurun_out/s17_launches_train.csv python tests/diag_train_steps.py > gpurun_out/s17_ncu_list.log 2>&1
STEPS=3 MODE=tf32x3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 152 -c 76 --csv --log-file gpurun_out/s17_launches_train_x3.csv python tests/diag_train_steps.py > gpurun_out/s17_ncu_list_x3.log 2>&1
for K in st_compose_steps st_bwd_staged st_wb_bwd_axis st_fwd_staged heads_bwd_k bce_loss_k lstm_bwd_k; do
  STEPS=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/s17_$K python tests/diag_train_steps.py > gpurun_out/s17_ncu_$K.log 2>&1
done
timeout 300 python bench.py --workload st > gpurun_out/s17_bench_st.json 2> gpurun_out/s17_bench_st.err
timeout 400 python examples/train_synthetic.py --iters 25000 --every 1000 --gemm tf32x3 --log gpurun_out/s17_gpu_convergence_tf32x3.log > /dev/null 2>&1
timeout 400 python examples/train_synthetic.py --iters 25000 --every 1000 --gemm tf32 --log gpurun_out/s17_gpu_convergence_tf32.log > /dev/null 2>&1
timeout 200 python tests/diag_loader.py 60000 > gpurun_out/s17_loader.txt 2>&1

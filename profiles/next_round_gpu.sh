#!/bin/bash
# One gpurun call that collects what round 1 could not measure any more (its GPU minutes were spent):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/next_round_gpu.sh'
# Everything lands in gpurun_out/ (scratch); copy the summaries you want judged into profiles/.
set -x
mkdir -p gpurun_out
# 1. the GPU tests written after the last GPU run (the reference-graph vectors, trained weights, image summary)
python -m pytest tests/test_gpu_zz_reference_graph.py -q -m gpu 2>&1 | tail -30 > gpurun_out/zz_tests.log
# 2. the whole GPU suite
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/gpu_tests.log
# 3. bench lines: train (default), ST microbench, inference, reference arm
python bench.py > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
python bench.py --workload st > gpurun_out/bench_st.json 2> gpurun_out/bench_st.err
python bench.py --workload infer > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# 4. ncu launch list of the FINAL 75-kernel train step (the committed r1_train_launches_e.md predates two ST changes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_train.csv "train step launch list, final schedule, TF32 mode, B=4096" \
    "ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv python bench.py --workload train --steps 3 --warmup 3" \
    > gpurun_out/train_launches.md 2>&1 || true
# 5. training through the CUDA path on synthetic canvases (counterpart of profiles/r1_oracle_convergence.log)
timeout 600 python examples/train_synthetic.py --iters 25000 --every 1000 --log gpurun_out/gpu_convergence.log \
    > gpurun_out/gpu_convergence.out 2>&1

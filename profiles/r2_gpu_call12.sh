set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -15 > gpurun_out/s12_tests.log
timeout 600 python bench.py > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err
MODE=tf32 timeout 200 python tests/diag_step_profile.py > gpurun_out/s12_step_profile_tf32.txt 2>&1
MODE=tf32x3 timeout 200 python tests/diag_step_profile.py > gpurun_out/s12_step_profile_x3.txt 2>&1

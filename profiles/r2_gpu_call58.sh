set -x
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
PART=st timeout 500 $S --tool memcheck python tests/sanitizer_smoke_r2.py > gpurun_out/s58_memcheck_st.txt 2>&1
PART=st timeout 500 $S --tool racecheck python tests/sanitizer_smoke_r2.py > gpurun_out/s58_racecheck_st.txt 2>&1
PART=st timeout 300 $S --tool synccheck python tests/sanitizer_smoke_r2.py > gpurun_out/s58_synccheck_st.txt 2>&1
PART=gemm AIR_TC_PERSIST=1 timeout 300 $S --tool memcheck python tests/sanitizer_smoke_r2.py > gpurun_out/s58_memcheck_gemm_persist.txt 2>&1
PART=gemm AIR_TC_CLUSTER4=1 timeout 300 $S --tool memcheck python tests/sanitizer_smoke_r2.py > gpurun_out/s58_memcheck_gemm_cluster4.txt 2>&1
PART=model timeout 500 $S --tool memcheck python tests/sanitizer_smoke_r2.py > gpurun_out/s58_memcheck_model.txt 2>&1

set -x
mkdir -p gpurun_out
{
for cfg in "" "AIR_TC_BN=64" "AIR_TC_BN=64 AIR_TC_STAGES=2" "AIR_TC_BN=64 AIR_TC_STAGES=3" "AIR_TC_STAGES=2" "AIR_TC_PERSIST=2" "AIR_TC_PERSIST=1"; do
  echo "== $cfg"; env $cfg TIME=1 SHAPE="fwd gm rng" timeout 100 python tests/diag_one_gemm.py
  env $cfg TIME=1 SHAPE="fwd gm" timeout 100 python tests/diag_one_gemm.py
done
} > gpurun_out/s49_gm_rng.txt 2>&1
AIR_PDL=0 timeout 200 python tests/diag_step_profile.py > gpurun_out/s49_step_kernels_warm.txt 2>&1
AIR_PDL=0 MODE=tf32x3 timeout 200 python tests/diag_step_profile.py > gpurun_out/s49_step_kernels_warm_x3.txt 2>&1

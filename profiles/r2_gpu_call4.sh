set -x
mkdir -p gpurun_out
AIR_TC_PAIR=256 MODE=tf32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair -s 4 -c 1 -o gpurun_out/s4_pair256_big python tests/diag_roofline_gemm.py > gpurun_out/s4_ncu.log 2>&1
for S in 2 3 4 6; do
AIR_TC_PAIR=256 AIR_TC_STAGES=$S MODE=tf32 timeout 120 python tests/diag_roofline_gemm.py > gpurun_out/s4_pair256_stages$S.txt 2>&1
done
AIR_TC_PAIR=128 AIR_TC_STAGES=4 MODE=tf32 timeout 120 python tests/diag_roofline_gemm.py > gpurun_out/s4_pair128_stages4.txt 2>&1

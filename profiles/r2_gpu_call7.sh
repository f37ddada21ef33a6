set -x
mkdir -p gpurun_out
{
for cfg in "AIR_TC_PAIR=0" "AIR_TC_PAIR=256" "AIR_TC_PAIR=256 AIR_PDL=0" "AIR_TC_PAIR=256 AIR_TC_STAGES=3" "AIR_TC_PAIR=256 AIR_TC_STAGES=3 AIR_PDL=0" "AIR_TC_PAIR=128" "AIR_TC_PAIR=256 MODE=tf32x3" "AIR_TC_PAIR=256 MODE=tf32x3 AIR_PDL=0"; do
  echo "== $cfg"
  env $cfg timeout 120 python tests/diag_pair_loop.py
done
} > gpurun_out/s7_pair_loop.txt 2>&1

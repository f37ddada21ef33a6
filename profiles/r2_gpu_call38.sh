set -x
mkdir -p gpurun_out /tmp/ncu
cap() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 ncu --set full --clock-control none -k regex:gemm_tf32 -s 4 -c 1 -o /tmp/ncu/$name python tests/diag_one_gemm.py > gpurun_out/s38_$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/s38_$name.csv 2>/dev/null
}
cap persist_dxgm "SHAPE=dX gm dsp"
cap nonpersist_dxgm "SHAPE=dX gm dsp" AIR_TC_PERSIST=0
cap fwd_r1 "SHAPE=fwd r1 softplus"
ls -la /tmp/ncu

"""The dominant GEMM of the train step ([4096,2500]x[2500,1024]) alone, for an `ncu --set full` capture:
MODE=tf32|tf32x3 ncu --set full --clock-control none -k "regex:gemm_tf32" -s 4 -c 1 -o out python tests/diag_roofline_gemm.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, bench_train
print(bench_train.gemm_roofline(4096, os.environ.get("MODE", "tf32"), bench.measured_peaks(), steps=6))

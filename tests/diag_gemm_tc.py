"""Diagnostic (not a pytest test): runs the tcgen05 GEMM over layouts / shapes and prints the
max abs error per case instead of stopping at the first failure.  python tests/diag_gemm_tc.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from air_b200 import ops  # noqa: E402
from tests.test_gpu_ops import _tf32_case  # noqa: E402

ok = True
for (M, N, Kd) in ((128, 128, 32), (128, 64, 32), (128, 128, 64), (128, 128, 128), (256, 256, 256), (130, 70, 36),
                   (4096, 1024, 256), (784, 512, 4096)):
    for tA in (False, True):
        for tB in (False, True):
            try:
                got, want = _tf32_case(M, N, Kd, tA, tB, seed=1)
                torch.cuda.synchronize()
                err = np.abs(got - want).max()
                bad = (got != want).mean()
                print(f"M={M} N={N} K={Kd} tA={int(tA)} tB={int(tB)} maxerr={err:.3g} mismatch_frac={bad:.4f} "
                      f"got[0,:4]={got[0, :4]} want[0,:4]={want[0, :4]}", flush=True)
                ok &= err == 0
            except Exception as e:  # noqa: BLE001
                print(f"M={M} N={N} K={Kd} tA={int(tA)} tB={int(tB)} EXC {type(e).__name__}: {e}", flush=True)
                ok = False
print("ALL OK" if ok else "SOME FAILED")
mode = ab._cabi.GEMM_MODES["tf32"]
for (M, N, Kd, tA, tB) in ((4096, 1024, 2500, False, False), (2500, 1024, 4096, True, False), (4096, 784, 512, False, False),
                           (4096, 512, 784, False, True), (784, 512, 4096, True, False), (4096, 1024, 256, False, False)):
    A = torch.randn((Kd, M) if tA else (M, Kd), device="cuda")
    B = torch.randn((N, Kd) if tB else (Kd, N), device="cuda")
    out = torch.empty(M, N, device="cuda")
    for md, name in ((mode, "tf32"), (0, "fp32")):
        try:
            for _ in range(3):
                ops.gemm(A, B, out, tA=tA, tB=tB, mode=md)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.gemm(A, B, out, tA=tA, tB=tB, mode=md)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"{name} M={M} N={N} K={Kd} tA={int(tA)} tB={int(tB)}: {ms * 1e3:.1f} us  {2.0 * M * N * Kd / ms / 1e9:.1f} TFLOP/s",
                  flush=True)
        except Exception as e:  # noqa: BLE001
            print(name, "EXC", e)

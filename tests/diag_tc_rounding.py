"""Diagnostic (not a pytest test): how does tcgen05.mma kind::tf32 round?  python tests/diag_tc_rounding.py

 1. operand conversion: is an fp32 operand truncated or rounded to TF32 by the tensor core?
 2. accumulation: with TF32-exact positive operands (every product exact in FP32, all partial sums positive) the only
    error is the rounding of the accumulate step; its SIGN tells round-to-nearest (mean ~ 0) from round-toward-zero
    (mean ~ -0.5 ulp per accumulation).
 3. the 3xTF32 mode with 1..3 hi*hi accumulator chains (AIR_TC_CHAINS is read once per process, so each setting runs
    in a subprocess): error vs fp64 next to the exact-FP32 FMA chain's.
"""
import os
import subprocess
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from air_b200 import ops  # noqa: E402

MODES = ab._cabi.GEMM_MODES


def gemm(A, B, mode, ws=None):
    out = torch.empty(A.shape[0], B.shape[1], device="cuda")
    ops.gemm(A, B, out, mode=MODES[mode], ws=ws)
    return out


def trunc_tf32(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def part1():
    A = torch.zeros(128, 32, device="cuda")
    B = torch.zeros(32, 64, device="cuda")
    vals = [1 + 2.0 ** -11 + 2.0 ** -13, 1 + 2.0 ** -10 - 2.0 ** -23, 1 + 2.0 ** -11, -(1 + 2.0 ** -11 + 2.0 ** -13)]
    for i, v in enumerate(vals):
        A[i, 0] = v
    B[0, :] = 1.0
    out = gemm(A, B, "tf32")[:len(vals), 0].cpu().double().numpy()
    for v, o in zip(vals, out):
        t = float(trunc_tf32(torch.tensor([v], dtype=torch.float32))[0])
        kind = "TRUNCATED" if o == t else ("rounded-to-nearest" if abs(o - v) <= 2.0 ** -11 else "?")
        print(f"operand {v!r}: tensor core used {o!r}  (trunc would be {t!r}) -> {kind}")


def part2():
    rng = np.random.RandomState(0)
    for Kd in (256, 1024, 4096):
        A = trunc_tf32(torch.from_numpy((1.0 + rng.rand(256, Kd)).astype(np.float32)))
        B = trunc_tf32(torch.from_numpy((1.0 + rng.rand(Kd, 128)).astype(np.float32)))
        want = A.double() @ B.double()
        ulp = np.spacing(want.float().numpy()).astype(np.float64)
        for mode in ("tf32", "fp32"):
            got = gemm(A.cuda(), B.cuda(), mode).cpu().double()
            e = (got - want).numpy() / ulp
            print(f"K={Kd:5d} {mode:5s}: positive-sum error in ulps of the result: mean {e.mean():+8.2f}  rms {np.sqrt((e ** 2).mean()):7.2f}  "
                  f"min {e.min():+8.1f} max {e.max():+8.1f}   ({Kd // 8} accumulate steps of 8 products)")


def part3():
    rng = np.random.RandomState(1)
    ws = torch.zeros(16 << 20, device="cuda")
    print("AIR_TC_CHAINS =", os.environ.get("AIR_TC_CHAINS", "(default)"))
    for (M, N, Kd, tA, what) in ((4096, 1024, 2500, False, "xK fwd"), (1024, 512, 784, False, "enc1"), (784, 512, 12288, True, "dW long K"),
                                 (2500, 1024, 4096, True, "dKx")):
        for dist in ("randn", "positive"):
            f = (lambda *s: rng.randn(*s)) if dist == "randn" else (lambda *s: 0.5 + rng.rand(*s))
            A = torch.from_numpy(f(M, Kd).astype(np.float32))
            B = torch.from_numpy(f(Kd, N).astype(np.float32))
            want = A.double() @ B.double()
            nrm = np.linalg.norm(want.numpy())
            res = {}
            for mode in ("tf32", "tf32x3", "fp32"):
                Ad = A.t().contiguous().cuda() if tA else A.cuda()
                out = torch.empty(M, N, device="cuda")
                ops.gemm(Ad, B.cuda(), out, tA=tA, mode=MODES[mode], ws=ws)
                d = out.cpu().double() - want
                res[mode] = (np.linalg.norm(d.numpy()) / nrm, float((d / want).mean()) if dist == "positive" else float("nan"))
            print(f"{what:10s} M={M} N={N} K={Kd} {dist:8s}: " +
                  "  ".join(f"{m}: rel {r[0]:.2e} bias {r[1]:+.2e}" for m, r in res.items()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "part3":
        part3()
    else:
        part1()
        part2()
        for ch in ("1", "2", "3"):
            env = dict(os.environ, AIR_TC_CHAINS=ch)
            subprocess.run([sys.executable, os.path.abspath(__file__), "part3"], env=env, check=False)

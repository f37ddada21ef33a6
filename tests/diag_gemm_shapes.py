"""Diagnostic: device time of every GEMM shape of one AIR train step (B=4096, T=3), TF32 mode,
timed as a CUDA graph of 20 back-to-back launches (no host overhead, warm L2)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from air_b200 import ops  # noqa: E402

K = ab._cabi
mode = K.GEMM_MODES[os.environ.get("MODE", "tf32")]
ws = torch.zeros(16 << 20, device="cuda")

B, TB = 4096, 3 * 4096
shapes = [  # name, M, N, K, tA, tB, cinit, bias, epi(aux)
    ("fwd xK", B, 1024, 2500, 0, 0, 0, 0, 0), ("fwd hKh+xk+b", B, 1024, 256, 0, 0, 1, 1, 0), ("fwd hid relu", B, 320, 256, 0, 0, 0, 1, 1),
    ("fwd r1 softplus", B, 512, 784, 0, 0, 0, 1, 2), ("fwd r2 softplus", B, 256, 512, 0, 0, 0, 1, 2), ("fwd ml", B, 100, 256, 0, 0, 0, 1, 0),
    ("fwd g1 softplus", B, 256, 50, 0, 0, 0, 1, 2), ("fwd g2 softplus", B, 512, 256, 0, 0, 0, 1, 2), ("fwd gm", B, 784, 512, 0, 0, 0, 1, 0),
    ("dX gm dsp", B, 512, 784, 0, 1, 0, 0, 4), ("dX g2 dsp", B, 256, 512, 0, 1, 0, 0, 4), ("dX g1", B, 50, 256, 0, 1, 0, 0, 0),
    ("dX ml dsp", B, 256, 100, 0, 1, 0, 0, 4), ("dX r2 dsp", B, 512, 256, 0, 1, 0, 0, 4), ("dX r1", B, 784, 512, 0, 1, 0, 0, 0),
    ("dX hid +dh", B, 256, 320, 0, 1, 1, 0, 0), ("dX Kh", B, 256, 1024, 0, 1, 0, 0, 0),
    ("dW gm", 512, 784, TB, 1, 0, 0, 0, 0), ("dW g2", 256, 512, TB, 1, 0, 0, 0, 0), ("dW g1", 50, 256, TB, 1, 0, 0, 0, 0),
    ("dW ml", 256, 100, TB, 1, 0, 0, 0, 0), ("dW r2", 512, 256, TB, 1, 0, 0, 0, 0), ("dW r1", 784, 512, TB, 1, 0, 0, 0, 0),
    ("dW hid", 256, 320, TB, 1, 0, 0, 0, 0), ("dW Kh", 256, 1024, 2 * B, 1, 0, 0, 0, 0), ("dW Kx", 2500, 1024, B, 1, 0, 0, 0, 0),
]
mult = {"fwd xK": 1, "dW Kx": 1}
if os.environ.get("M_FULL", "0") != "0":   # the VAE GEMMs as the model runs them: once per step on T*B rows
    vae = ("r1", "r2", "ml", "g1", "g2", "gm")
    shapes = [(n, TB if (n.split()[1] in vae and not n.startswith("dW")) else M, N, Kd, tA, tB, ci, bi, epi)
              for n, M, N, Kd, tA, tB, ci, bi, epi in shapes]
    mult.update({n: 1 for n, *_ in shapes if n.split()[1] in vae})
tot = 0.0
print(f"{'gemm':18s} {'M':>6s} {'N':>5s} {'K':>6s}   us     TFLOP/s  x/step  us/step")
for name, M, N, Kd, tA, tB, ci, bi, epi in shapes:
    pad = lambda n: (n + 3) // 4 * 4
    A = torch.randn((Kd, pad(M)) if tA else (M, pad(Kd)), device="cuda")[:, :(M if tA else Kd)]
    Bm = torch.randn((N, pad(Kd)) if tB else (Kd, pad(N)), device="cuda")[:, :(Kd if tB else N)]
    out = torch.empty(M, N, device="cuda")
    Cinit = torch.randn(M, N, device="cuda") if ci else None
    bias = torch.randn(N, device="cuda") if bi else None
    aux = torch.rand(M, N, device="cuda") if epi in (3, 4) else None
    run = lambda: ops.gemm(A, Bm, out, Cinit=Cinit, bias=bias, aux=aux, tA=bool(tA), tB=bool(tB), epi=epi, mode=mode, ws=ws)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            run()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 100
    n = mult.get(name, 3 if not name.startswith("dW") else 1)
    tot += us * n
    print(f"{name:18s} {M:6d} {N:5d} {Kd:6d} {us:7.1f} {2.0 * M * N * Kd / us / 1e6:9.1f} {n:6d} {us * n:8.1f}")
print(f"total GEMM us/step: {tot:.1f}")

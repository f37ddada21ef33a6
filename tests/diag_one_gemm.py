"""Diagnostic: ONE GEMM shape of the train step, a few launches (for ncu captures).
SHAPE="name" (a row of tests/diag_gemm_shapes.py, default "dX gm dsp") MODE=tf32|tf32x3 M_ROWS=12288"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from air_b200 import ops  # noqa: E402

SHAPES = {  # name: N, K, tA, tB, cinit, bias, epi
    "fwd r1 softplus": (512, 784, 0, 0, 0, 1, 2), "fwd gm": (784, 512, 0, 0, 0, 1, 0), "dX gm dsp": (512, 784, 0, 1, 0, 0, 4),
    "dX r1": (784, 512, 0, 1, 0, 0, 0), "fwd r2 softplus": (256, 512, 0, 0, 0, 1, 2), "fwd hKh": (1024, 256, 0, 0, 1, 1, 0),
    "fwd gm rng": (784, 512, 0, 0, 0, 1, 6),   # the gen_mean layer as the model runs it: noise generated in the epilogue
}
name = os.environ.get("SHAPE", "dX gm dsp")
N, Kd, tA, tB, ci, bi, epi = SHAPES[name]
M = int(os.environ.get("M_ROWS", "12288"))
mode = ab._cabi.GEMM_MODES[os.environ.get("MODE", "tf32")]
ws = torch.zeros(16 << 20, device="cuda")
A = torch.randn(M, Kd, device="cuda")
Bm = torch.randn((N, Kd) if tB else (Kd, N), device="cuda")
out = torch.empty(M, N, device="cuda")
Cinit = torch.randn(M, N, device="cuda") if ci else None
bias = torch.randn(N, device="cuda") if bi else None
aux = torch.rand(M, N, device="cuda") if epi in (3, 4) else ops.rng_state(1, "cuda") if epi == 6 else None
run = lambda: ops.gemm(A, Bm, out, Cinit=Cinit, bias=bias, aux=aux, tA=bool(tA), tB=bool(tB), epi=epi, mode=mode, ws=ws, epi_param=0.3)
for _ in range(6):
    run()
torch.cuda.synchronize()
if os.environ.get("TIME", "0") != "0":
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
    print("us per launch: %.1f" % (e0.elapsed_time(e1) * 10.0))
print(name, M, N, Kd, "done")

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab
from air_b200 import ops
ws = torch.zeros(16 << 20, device="cuda"); 
def t(run):
    for _ in range(3): run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): run()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 10
M = 4096
for (N, Kd, lda, ldb_pad, tB) in [(512, 784, 784, 0, 0), (512, 784, 800, 0, 0), (512, 784, 1024, 0, 0), (512, 768, 768, 0, 0),
                                   (512, 784, 784, 0, 1), (512, 784, 800, 16, 1), (784, 512, 512, 0, 0), (512, 800, 800, 0, 0)]:
    A = torch.randn(M, lda, device="cuda")[:, :Kd]
    if tB: Bm = torch.randn(N, Kd + ldb_pad, device="cuda")[:, :Kd]
    else: Bm = torch.randn(Kd, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    us = t(lambda: ops.gemm(A, Bm, out, tB=bool(tB), mode=1))
    print(f"N={N} K={Kd} lda={lda} tB={tB} ldb_pad={ldb_pad}: {us:.1f} us  {2.0*M*N*Kd/us/1e6:.0f} TFLOP/s")

import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab
from air_b200 import ops
M, N, Kd = 4096, 512, 256
A = torch.randn(M, Kd, device="cuda"); W = torch.randn(Kd, N, device="cuda"); b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
for _ in range(6):
    ops.gemm(A, W, out, bias=b, epi=2, mode=1)
torch.cuda.synchronize()

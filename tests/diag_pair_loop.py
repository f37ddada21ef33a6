"""Diagnostic: the big GEMM ([4096,2500]x[2500,1024]) launched back to back, eagerly and from a CUDA graph, with a
different output buffer per launch or the same one; prints us per launch.  Env: MODE, AIR_TC_PAIR, AIR_PDL, AIR_TC_STAGES."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from air_b200 import ops  # noqa: E402

mode = ab._cabi.GEMM_MODES[os.environ.get("MODE", "tf32")]
M, N, K = [int(v) for v in os.environ.get("SHAPE", "4096,1024,2500").split(",")]
x = torch.rand(M, K, device="cuda")
W = torch.randn(K, N, device="cuda") * 0.02
outs = [torch.empty(M, N, device="cuda") for _ in range(20)]


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def loop(n, same):
    for i in range(n):
        ops.gemm(x, W, outs[0 if same else i % 20], mode=mode)


for same in (True, False):
    eager = timed(lambda: loop(20, same), 5) / 20
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loop(20, same)
    graph = timed(g.replay, 5) / 20
    single = []
    for _ in range(5):                       # one launch at a time, device idle before and after
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(x, W, outs[0], mode=mode)
        e1.record()
        torch.cuda.synchronize()
        single.append(e0.elapsed_time(e1) * 1e3)
    print(f"same_out={same}: eager x20 {eager:.1f} us/launch, graph x20 {graph:.1f} us/launch, isolated launches {min(single):.1f}..{max(single):.1f} us",
          flush=True)

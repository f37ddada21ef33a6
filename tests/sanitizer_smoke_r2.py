"""compute-sanitizer run for the kernels added in round 2: st_wb_bwd_ref (every branch: clipped / mirrored / outside /
sheared windows, all flag combinations, single-step and batched-over-T entry points), the row-wise st_compose_steps
(T = 3, 5 and the generic fallback), the persistent and the 2 x 2 cluster GEMM (forced through AIR_TC_PERSIST /
AIR_TC_CLUSTER4 by the caller), air_adam_step_ex, the conv kernels for 4 and 16 filters, one model step.
    compute-sanitizer --tool memcheck|racecheck|synccheck python tests/sanitizer_smoke_r2.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from air_b200 import ops  # noqa: E402
from tests.test_gpu_st import _axis_cases  # noqa: E402

dev = "cuda"
part = os.environ.get("PART", "all")
if part in ("all", "st"):
    thi = torch.from_numpy(_axis_cases(np.random.RandomState(43))).to(dev)
    n = thi.shape[0]
    win, z, g = torch.rand(n, 28, 28, device=dev), torch.rand(n, device=dev), torch.randn(n, 50, 50, device=dev) * 1e6
    stop = (torch.rand(n, device=dev) > 0.8).float() * 1.5
    dw, dt, dz = torch.empty_like(win), torch.empty(n, 6, device=dev), torch.empty(n, device=dev)
    for flags in (4, 5, 6, 7):
        ops.writeback_canvas_bwd(win, thi, z, stop, 0.99, g, dw, dt, dz, 28, 28, 50, 50, window_is_sigmoid=bool(flags & 1),
                                 axis_aligned_theta=bool(flags & 2), reference_rounding=True)
    for T in (3, 5, 2):
        B = 6
        wins = torch.rand(T, B, 784, device=dev)
        th = thi[:T * B].reshape(T, B, 6).contiguous()
        fields = torch.zeros(T, 2, B, device=dev)
        fields[:, 0] = torch.rand(T, B, device=dev)
        fields[:, 1] = (torch.rand(T, B, device=dev) > 0.7).float() * 1.5
        canvas, dcanvas = torch.empty(B, 2500, device=dev), torch.randn(B, 2500, device=dev)
        ops.writeback_canvas_fwd_steps(wins, th, fields[0, 0], fields[0, 1], 2 * B, 0.99, None, canvas, 28, 28, 50, 50)
        dgen, dthi, dzz = torch.empty(T, B, 784, device=dev), torch.empty(T, B, 6, device=dev), torch.empty(T, B, device=dev)
        ops.writeback_canvas_bwd_steps(wins, th, fields[0, 0], fields[0, 1], 2 * B, 0.99, dcanvas, dgen, dthi, dzz, 28, 28, 50, 50,
                                       window_is_sigmoid=True, axis_aligned_theta=True, reference_rounding=True)
    torch.cuda.synchronize()
    print("st part done")
if part in ("all", "gemm"):
    ws = torch.zeros(4 << 20, device=dev)
    for M, N, K, tA, tB in ((640, 384, 200, False, False), (1000, 520, 136, False, True), (512, 256, 96, True, False), (300, 140, 64, True, True)):
        A = torch.randn((K, M) if tA else (M, K), device=dev)
        Bm = torch.randn((N, K) if tB else (K, N), device=dev)
        out = torch.empty(M, N, device=dev)
        ops.gemm(A, Bm, out, bias=torch.randn(N, device=dev), tA=tA, tB=tB, epi=1, mode=ab._cabi.GEMM_MODES["tf32"], ws=ws)
    torch.cuda.synchronize()
    print("gemm part done", os.environ.get("AIR_TC_PERSIST"), os.environ.get("AIR_TC_CLUSTER4"))
if part in ("all", "model"):
    from tests.parity_util import cuda_noise, make_pair, realistic_fixture  # noqa: E402
    imgs, cnt, params, noise = realistic_fixture(8, seed=1)
    _, m = make_pair(imgs, cnt, params, train=True, gemm_mode="tf32")
    m.skip_nonfinite_updates = True
    m.train_step(cuda_noise(noise))
    m.train_step()
    for F_ in (4, 16):
        ab.reset_variable_scopes()
        mc = ab.AIRModel(imgs.cuda(), cnt.cuda(), train=True, cnn=True, cnn_filters=F_, gemm_mode="tf32", scope=f"c{F_}")
        mc.train_step()
    torch.cuda.synchronize()
    print("model part done")

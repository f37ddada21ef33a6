"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): every kernel of the
library on tiny shapes, both GEMM modes.  python tests/sanitizer_smoke.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from oracle import air_oracle as O  # noqa: E402
from tests.parity_util import covered_fixture, cuda_noise, make_pair, realistic_fixture  # noqa: E402

for mode in ("fp32", "tf32"):
    for fx in (covered_fixture, realistic_fixture):
        imgs, cnt, params, noise = fx(8, seed=1)
        _, m = make_pair(imgs, cnt, params, train=True, gemm_mode=mode)
        m.train_step(cuda_noise(noise))
        m.train_step()
        torch.cuda.synchronize()
        print(mode, fx.__name__, "loss", m.loss.item())
# rotated / multi-channel / generic ST paths, fused canvas, concrete ops
U = torch.rand(5, 12, 12, 1, device="cuda", requires_grad=True)
th = (torch.tensor([[0.8, 0.2, 0.0, -0.2, 0.7, 0.1]], device="cuda") + 0.05 * torch.randn(5, 6, device="cuda")).requires_grad_(True)
ab.transformer(U, th, (6, 8)).sum().backward()
U3 = torch.rand(3, 9, 7, 3, device="cuda", requires_grad=True)
th3 = torch.randn(3, 6, device="cuda", requires_grad=True)
ab.transformer(U3, th3, (5, 4)).sum().backward()
# warp-specialised write-back backward: every branch (two row blocks, tiny / mirrored / clipped / outside windows, shear)
import numpy as np  # noqa: E402
from tests.test_gpu_st import _axis_cases  # noqa: E402
thi = torch.from_numpy(_axis_cases(np.random.RandomState(43))).cuda()
n = thi.shape[0]
win, z, g = torch.rand(n, 28, 28, device="cuda"), torch.rand(n, device="cuda"), torch.randn(n, 50, 50, device="cuda")
stop = (torch.rand(n, device="cuda") > 0.8).float() * 1.5
dw, dt, dz = torch.empty_like(win), torch.empty(n, 6, device="cuda"), torch.empty(n, device="cuda")
for flags in (0, 1, 2, 3):
    ab.ops.writeback_canvas_bwd(win, thi, z, stop, 0.99, g, dw, dt, dz, 28, 28, 50, 50, window_is_sigmoid=bool(flags & 1),
                                axis_aligned_theta=bool(flags & 2))
torch.cuda.synchronize()
print("sanitizer smoke done")

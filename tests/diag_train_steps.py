"""A few EAGER train steps (B = 4096, T = 3) for ncu captures: every kernel is its own launch (no CUDA graph).
MODE=tf32|tf32x3 STEPS=3 ncu ... python tests/diag_train_steps.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench_train  # noqa: E402

m, imgs, cnt, data = bench_train._model(4096, os.environ.get("MODE", "tf32"), seed=0, train=True, max_steps=3)
import air_b200 as ab  # noqa: E402
for _ in range(int(os.environ.get("STEPS", "3"))):
    n0 = ab.launch_count()
    m.train_step()
    per_step = ab.launch_count() - n0
torch.cuda.synchronize()
print("loss", float(m.loss), "kernels_per_step", per_step)

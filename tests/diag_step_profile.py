"""Warm per-kernel timing of the eager train step with torch.profiler (CUPTI), B = 4096, TF32 mode.
usage (GPU box): python tests/diag_step_profile.py > gpurun_out/step_profile.txt"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_train

m, imgs, cnt, data = bench_train._model(4096, os.environ.get("MODE", "tf32"), seed=0, train=True, max_steps=3, cnn=bool(int(os.environ.get("CNN", "0"))))
for _ in range(5):
    m.train_step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        m.train_step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None: t = e.cuda_time_total
    if t > 0: rows.append((t / N, e.count / N, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("eager step: %.1f us of kernel time per step, %d kernel names" % (tot, len(rows)))
for t, c, k in rows:
    print("%8.1f us  %5.1f%%  x%-5.1f  avg %7.2f  %s" % (t, 100 * t / tot, c, t / c, k[:110]))

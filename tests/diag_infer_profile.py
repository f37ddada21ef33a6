"""Warm per-kernel timing of the inference step (train=False, B = 65536, T = 5, TF32) with torch.profiler."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_train
m, imgs, cnt, data = bench_train._model(65536, "tf32", seed=0, train=False, max_steps=5)
for _ in range(3):
    m.run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        m.run()
    torch.cuda.synchronize()
rows = [(getattr(e, "device_time_total", 0) / N, e.count / N, e.key) for e in prof.key_averages() if getattr(e, "device_time_total", 0) > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("inference step: %.1f us of kernel time per step" % tot)
for t, c, k in rows[:24]:
    print("%8.1f us  %5.1f%%  x%-5.1f  avg %7.2f  %s" % (t, 100 * t / tot, c, t / c, k[:100]))

"""Diagnostic: the first few hundred optimisation steps of the CUDA model and of the CPU oracle from the same default
initialisation, on the same batches with the same injected noise (reference training configuration, batch 64): prints
both loss curves.  Variants: VARIANT=rng (device-generated noise, graph replay) | inject (torch noise, eager)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from oracle import air_oracle as O  # noqa: E402

iters = int(os.environ.get("ITERS", "300"))
mode = os.environ.get("MODE", "tf32x3")
B = 64
train, cnt = ab.data.device_canvases(20000, seed=0)
train_c, cnt_c = train.cpu(), cnt.cpu()
params = O.init_params(seed=0)
orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING, train=True)
ab.reset_variable_scopes()
m = ab.AIRModel(train[:B].clone(), cnt[:B].clone(), train=True, annealing_schedules=O.DEFAULT_ANNEALING, gemm_mode=mode,
                **O.DEFAULT_HYPER)
m.store.load_named({k: v.cuda() for k, v in params.items()})
m2 = ab.AIRModel(train[:B].clone(), cnt[:B].clone(), train=True, annealing_schedules=O.DEFAULT_ANNEALING, gemm_mode=mode,
                 scope="rng", **O.DEFAULT_HYPER)
m2.store.load_named({k: v.cuda() for k, v in params.items()})
m2.capture()
g = torch.Generator().manual_seed(1)
lo, lg, lr = [], [], []
for it in range(iters):
    idx = torch.randint(0, 20000, (B,), generator=g)
    noise = O.make_noise(it, 3, B)
    if it < int(os.environ.get("ORACLE_ITERS", "150")):
        out, _ = orc.train_step(train_c[idx], cnt_c[idx], noise)
        lo.append(float(out["loss"]))
    m.feed(train[idx.cuda()], cnt[idx.cuda()])
    m.train_step({k: v.cuda() for k, v in noise.items()})
    lg.append(float(m.loss))
    m2.feed(train[idx.cuda()], cnt[idx.cuda()])
    m2.train_step()
    lr.append(float(m2.loss))
    if it % 25 == 24:
        a = lambda v: sum(v[-25:]) / 25 if len(v) >= it + 1 else float("nan")
        print(f"it {it + 1:5d}  oracle {a(lo):9.2f}   cuda(injected noise, eager) {a(lg):9.2f}   cuda(device noise, graph) {a(lr):9.2f}   "
              f"|g| {float(m.store.state[3]):.3e}", flush=True)
d = {k: float((v.cpu() - orc.params[k]).norm() / (orc.params[k] - params[k]).norm().clamp_min(1e-30)) for k, v in m.store.named_views().items()} \
    if len(lo) == iters else {}
if d:
    print("relative parameter-movement difference vs oracle after", iters, "steps:", {k: round(v, 3) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:6]})

"""Diagnostic: the CUDA step teacher-forced along the ORACLE's own training trajectory.  Every EVERY iterations the
oracle's current parameters, Adam slots, beta powers and global step are loaded into the CUDA model and one optimisation
step with the same batch and the same injected noise is compared tensor by tensor: loss, the 36 gradients, the global
norm and the parameter update.  A step whose gradients agree but whose update differs points at clip/Adam; a tensor whose
gradient departs only late in training points at a kernel path the initialisation never exercises.
ITERS=300 EVERY=10 MODE=tf32x3|fp32 python tests/diag_teacher_forced.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from oracle import air_oracle as O  # noqa: E402
from tests.parity_util import relnorm  # noqa: E402

iters = int(os.environ.get("ITERS", "300"))
every = int(os.environ.get("EVERY", "10"))
mode = os.environ.get("MODE", "tf32x3")
B = int(os.environ.get("BATCH", "64"))
torch.set_num_threads(int(os.environ.get("THREADS", "16")))
train, cnt = ab.data.device_canvases(20000, seed=0)
train_c, cnt_c = train.cpu(), cnt.cpu()
params = O.init_params(seed=0)
orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING, train=True)
ab.reset_variable_scopes()
m = ab.AIRModel(train[:B].clone(), cnt[:B].clone(), train=True, annealing_schedules=O.DEFAULT_ANNEALING, gemm_mode=mode,
                **O.DEFAULT_HYPER)
g = torch.Generator().manual_seed(1)


def load_state():
    st = m.store
    st.load_named({k: v.cuda() for k, v in orc.params.items()})
    am, av = st.named_adam()
    for k in orc.params:
        am[k].copy_(orc.adam_m[k].cuda())
        av[k].copy_(orc.adam_v[k].cuda())
    st.state[0] = float(orc.beta1_power)
    st.state[1] = float(orc.beta2_power)
    st.global_step = orc.global_step


for it in range(iters):
    idx = torch.randint(0, 20000, (B,), generator=g)
    noise = O.make_noise(it, 3, B)
    x, c = train_c[idx], cnt_c[idx]
    check = it % every == 0 or it < 3
    if check:
        before = {k: v.clone() for k, v in orc.params.items()}
        load_state()
        m.feed(train[idx.cuda()], cnt[idx.cuda()])
        m.train_step({k: v.cuda() for k, v in noise.items()})
        gl = float(m.loss)
        gg = {k: v.detach().cpu().clone() for k, v in m.store.named_grads().items()}
        gp = {k: v.detach().cpu().clone() for k, v in m.store.named_views().items()}
        gnorm = float(m.store.state[3])
        digits = m.rec_num_digits.cpu().clone()
        zsum = float(m.z_pres.sum())
        scales = m.rec_scales.cpu().clone()
    # the oracle's own step (unclipped gradients are recomputed for the comparison)
    if check:
        out_u, grads_u = orc.loss_and_grads(x, c, noise)
    out, grads = orc.train_step(x, c, noise)
    if check:
        ge = {k: relnorm(gg[k], grads_u[k]) for k in grads_u}
        worst = sorted(ge.items(), key=lambda kv: -kv[1])[:4]
        tot = relnorm(torch.cat([gg[k].reshape(-1) for k in grads_u]), torch.cat([grads_u[k].reshape(-1) for k in grads_u]))
        ue = {k: relnorm(gp[k] - before[k], orc.params[k] - before[k]) for k in grads_u}
        uworst = sorted(ue.items(), key=lambda kv: -kv[1])[:3]
        dig_eq = float((digits == out["rec_num_digits"].to(digits.dtype)).float().mean()) if "rec_num_digits" in out else -1
        print(f"it {it:4d} loss oracle {float(out['loss']):10.3f} cuda {gl:10.3f} | |g| oracle {float(out['grad_global_norm']):9.3e} "
              f"cuda {gnorm:9.3e} | grad rel err all {tot:8.2e} worst {[(k, round(v, 5)) for k, v in worst]} | "
              f"update rel err worst {[(k, round(v, 5)) for k, v in uworst]} | digits equal {dig_eq:.3f} | "
              f"mean scale {float(scales.mean()):.3f} z_sum {zsum:.1f}", flush=True)

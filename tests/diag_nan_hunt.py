"""Diagnostic: find the first optimisation step of a training run (examples/train_synthetic.py, SEED) whose loss is not
finite and report which buffers of that step are not finite / extreme.  Deterministic replay: state is check-pointed every
EVERY steps; on a NaN the last check-point is restored and the steps are repeated one at a time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402

seed = int(os.environ.get("SEED", "7"))
every = int(os.environ.get("EVERY", "50"))
iters = int(os.environ.get("ITERS", "6000"))
data = ab.data
train, cnt = data.device_canvases(60000, seed=0)
ab.reset_variable_scopes()
m = ab.AIRModel(train[:64].clone(), cnt[:64].clone(), train=True, annealing_schedules=data.TRAINING_ANNEALING,
                gemm_mode="tf32x3", seed=seed, **data.TRAINING_HYPER)
if os.environ.get("CAPTURE", "0") != "0":
    if os.environ.get("WITH_EVAL", "0") != "0":   # as examples/train_synthetic.py: a test-mode model sharing the variables
        val, vcnt = data.device_canvases(4096, seed=12345)
        ev = ab.AIRModel(val, vcnt, train=False, reuse=True, annealing_schedules=data.TRAINING_ANNEALING, gemm_mode="tf32x3",
                         **data.TRAINING_HYPER)
    m.capture()
g = torch.Generator(device="cuda").manual_seed(1 + seed)
st = m.store


def snapshot():
    return dict(flat=st.flat.clone(), m=st.adam_m.clone(), v=st.adam_v.clone(), state=st.state.clone(),
                rng=m.w["rng_state"].clone(), gen=g.get_state())


def restore(s):
    st.flat.copy_(s["flat"]); st.adam_m.copy_(s["m"]); st.adam_v.copy_(s["v"]); st.state.copy_(s["state"])
    m.w["rng_state"].copy_(s["rng"]); g.set_state(s["gen"])


def step():
    idx = torch.randint(0, 60000, (64,), generator=g, device="cuda")
    m.feed(train[idx], cnt[idx])
    m.train_step()


def report(it):
    w = m.w
    print(f"--- first non-finite loss at iteration {it}: loss {float(m.loss)}")
    F = w["fields"]
    names = {"scale s": F[:, ab._cabi.F_S], "x": F[:, ab._cabi.F_X], "y": F[:, ab._cabi.F_Y], "z": F[:, ab._cabi.F_Z],
             "theta_inv": w["theta_inv"], "theta": w["theta"], "window in": w["win"], "ml": w["ml"], "recon": w["recon"],
             "canvas": w["canvas"], "loss acc": w["loss"], "rec_loss": w["rec_loss"], "dcanvas": w["dcanvas"],
             "dgen": w["vae_d"]["dgen"], "dtheta_inv": w["dtheta_inv"], "dz": w["dz"], "dwin": w["dwin"], "dtheta": w["dtheta"],
             "grad": st.grad, "params": st.flat, "hh": w["hh"], "h": w["h"], "gates": w["gates"]}
    for k, t in names.items():
        t = t.float()
        fin = torch.isfinite(t)
        print(f"  {k:12s} finite {int(fin.sum()):9d}/{t.numel():9d}  absmax(finite) {float(t[fin].abs().max()) if fin.any() else float('nan'):.4e}"
              f"  min {float(t[fin].min()) if fin.any() else float('nan'):.4e}")
    bad = ~torch.isfinite(w["canvas"]).all(dim=1)
    if bad.any():
        b = int(bad.nonzero()[0])
        print("  first image with a non-finite canvas:", b, "theta_inv per step:", w["theta_inv"][:, b].tolist(), "s:", F[:, ab._cabi.F_S, b].tolist(),
              "z:", F[:, ab._cabi.F_Z, b].tolist())
    ng = {k: int((~torch.isfinite(v)).sum()) for k, v in st.named_grads().items()}
    print("  non-finite gradient entries per tensor:", {k: v for k, v in ng.items() if v})


snap, snap_it = snapshot(), 0
it = 0
while it < iters:
    step()
    it += 1
    if it % every == 0:
        if not bool(torch.isfinite(m.loss)) or not bool(torch.isfinite(st.flat).all()):
            restore(snap)
            it = snap_it
            while True:
                pre = st.flat.clone()
                step()
                it += 1
                if not bool(torch.isfinite(m.loss)) or not bool(torch.isfinite(st.flat).all()):
                    print("parameters before the step finite:", bool(torch.isfinite(pre).all()), " after:", bool(torch.isfinite(st.flat).all()))
                    report(it)
                    sys.exit(0)
        snap, snap_it = snapshot(), it
print("no non-finite loss in", iters, "iterations")

"""Diagnostic (not a pytest test): localises gradient mismatches between the CUDA backward
schedule and oracle autograd by comparing d(loss)/d(intermediate) per step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import air_b200 as ab  # noqa: E402
from oracle import air_oracle as O  # noqa: E402
from tests.parity_util import covered_fixture, cuda_noise, default_fixture, make_pair, relnorm  # noqa: E402

B = 64
imgs, cnt, params, noise = covered_fixture(B, seed=3)
orc, m = make_pair(imgs, cnt, params, train=True)
leaves = {k: v.detach().clone().requires_grad_(True) for k, v in orc.params.items()}
orc.params = leaves
taps = {}
out = orc.forward(imgs, cnt, noise, taps=taps)
out["canvas_raw"].retain_grad()
out["loss"].backward()
m._debug = {}
m.loss_and_grads(cuda_noise(noise))
print("loss", m.loss.item(), out["loss"].item())
vd = m.w["vae_d"]
gc, oc = m.w["dcanvas"].cpu(), out["canvas_raw"].grad
print("dcanvas relnorm", relnorm(gc, oc), "canvas relnorm", relnorm(m.w["canvas"], out["canvas_raw"]))
per_img = ((gc - oc).norm(dim=1) / oc.norm(dim=1))
print("dcanvas per-image rel err: max", per_img.max().item(), "argmax", per_img.argmax().item(), "n>1e-4:", int((per_img > 1e-4).sum()))
bad = per_img.argmax().item()
diff = (gc[bad] - oc[bad]).abs()
idx = diff.topk(5).indices
print(" worst pixels", idx.tolist(), "gpu", gc[bad][idx].tolist(), "oracle", oc[bad][idx].tolist(),
      "canvas gpu", m.w["canvas"][bad].cpu()[idx].tolist(), "canvas oracle", out["canvas_raw"][bad][idx].tolist(), "x", imgs[bad][idx].tolist())
rec0 = taps["windows"][0]
e = ((vd["dgen"][0].cpu() - rec0.grad * rec0.detach() * (1 - rec0.detach())).norm(dim=1) / (rec0.grad * rec0.detach() * (1 - rec0.detach())).norm(dim=1))
print("dgen[0] per-image rel err: max", e.max().item(), "median", e.median().item(), "n>1e-3:", int((e > 1e-3).sum()))
for t in range(3):
    d = m._debug[t]
    rec = taps["windows"][t]
    print(f"--- step {t}")
    print(" dz        ", relnorm(d["dz"], taps["z_pres"][t].grad))
    print(" dtheta_inv", relnorm(d["dtheta_inv"].reshape(B, 2, 3), taps["st_backward"][t].grad))
    print(" dgen      ", relnorm(vd["dgen"][t], rec.grad * rec.detach() * (1 - rec.detach())))
    print(" dwin      ", relnorm(d["dwin"], taps["windows_in"][t].grad))
    print(" dtheta    ", relnorm(d["dtheta"].reshape(B, 2, 3), taps["thetas"][t].grad))
    gs = taps["scales"][t].grad[:, 0]
    gsh = taps["shifts"][t].grad
    # ds, dx, dy implied by the CUDA dtheta / dtheta_inv (same formula as heads_bwd)
    s = taps["scales"][t].detach()[:, 0]; x = taps["shifts"][t].detach()[:, 0]; y = taps["shifts"][t].detach()[:, 1]
    dt, di = d["dtheta"].cpu(), d["dtheta_inv"].cpu()
    ds = (dt[:, 0] + dt[:, 4]) - (di[:, 0] + di[:, 4]) / s ** 2 + (di[:, 2] * x + di[:, 5] * y) / s ** 2
    print(" ds (from cuda dtheta) vs oracle", relnorm(ds, gs), " dx", relnorm(dt[:, 2] - di[:, 2] / s, gsh[:, 0]))
print("--- parameter gradients")
for k, g in m.store.named_grads().items():
    print(f" {k:40s} {relnorm(g, leaves[k].grad):.3e}")

"""Generates the committed golden vectors under tests/golden/ from the CPU oracle.

The reference (TensorFlow 1.3) cannot run here and ships no vectors; these fixtures are outputs of
oracle/air_oracle.py (itself pinned to the reference's graph by make_golden_ref_graph.py), cross-checked bit-for-bit against the
independent C restatement oracle/st_oracle.c at generation time.  They pin the oracle
(regression) and give the GPU tests inputs/outputs that do not depend on /root/reference
or on re-running the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import air_oracle as O  # noqa: E402
from oracle import c_oracle as C  # noqa: E402


def air_thetas(rng, B):
    s = rng.uniform(0.3, 0.9, B).astype(np.float32)
    x = rng.uniform(-0.5, 0.5, B).astype(np.float32)
    y = rng.uniform(-0.5, 0.5, B).astype(np.float32)
    th = np.zeros((B, 2, 3), np.float32)
    th[:, 0, 0] = s; th[:, 1, 1] = s; th[:, 0, 2] = x; th[:, 1, 2] = y
    one = np.float32(1.0)
    thi = np.zeros((B, 2, 3), np.float32)
    thi[:, 0, 0] = one / s; thi[:, 1, 1] = one / s; thi[:, 0, 2] = -x / s; thi[:, 1, 2] = -y / s
    return th, thi


def main():
    rng = np.random.RandomState(1)
    B = 6
    # ---- ST crop 50x50 -> 28x28 and inverse 28x28 -> 50x50 (AIR shapes)
    imgs, _ = O.synthetic_canvases(B, seed=3)
    U = imgs.numpy().reshape(B, 50, 50, 1)
    th, thi = air_thetas(rng, B)
    crop = O.transformer(torch.from_numpy(U), torch.from_numpy(th), (28, 28)).numpy()
    assert np.array_equal(crop, C.st_forward(U, th, (28, 28)))
    win = rng.rand(B, 28, 28, 1).astype(np.float32)
    back = O.transformer(torch.from_numpy(win), torch.from_numpy(thi), (50, 50)).numpy()
    assert np.array_equal(back, C.st_forward(win, thi, (50, 50)))
    # fused canvas
    z = rng.rand(B).astype(np.float32)
    stop = np.array([0.1, 0.995, 0.5, 1.3, 0.98999, 0.99], np.float32)
    canvas = rng.rand(B, 2500).astype(np.float32)
    canvas_out = C.canvas_update(canvas, back.reshape(B, 2500), z, stop, 0.99)
    # general theta (rotation/shear), C = 3, odd sizes
    Ug = rng.rand(3, 11, 13, 3).astype(np.float32)
    thg = rng.uniform(-1.2, 1.2, (3, 2, 3)).astype(np.float32)
    outg = O.transformer(torch.from_numpy(Ug), torch.from_numpy(thg), (7, 9)).numpy()
    assert np.array_equal(outg, C.st_forward(Ug, thg, (7, 9)))
    # backward (fp32 autograd of the op-for-op oracle)
    Ut = torch.from_numpy(win).requires_grad_(True)
    tt = torch.from_numpy(thi).requires_grad_(True)
    g = torch.from_numpy(rng.randn(B, 50, 50, 1).astype(np.float32))
    O.transformer(Ut, tt, (50, 50)).backward(g)
    np.savez_compressed(os.path.join(HERE, "st.npz"), U=U, theta=th, theta_inv=thi, crop=crop, window=win, back=back,
                        z=z, stop=stop, canvas=canvas, canvas_out=canvas_out, Ug=Ug, thg=thg, outg=outg,
                        back_dout=g.numpy(), back_dU=Ut.grad.numpy(), back_dtheta=tt.grad.numpy())

    # ---- Concrete / ACT step
    n = 64
    lo = rng.randn(n).astype(np.float32) * 2
    u = rng.rand(n).astype(np.float32)
    u[:3] = [0.0, 1e-7, 0.9999999]
    stop_prev = rng.choice([0.0, 0.3, 0.98, 0.99, 1.5], n).astype(np.float32)
    loss_prev = rng.rand(n).astype(np.float32)
    dig_prev = rng.randint(0, 3, n).astype(np.int32)
    out = {}
    for train in (0, 1):
        for k, v in C.concrete_step(lo, u, stop_prev, loss_prev, dig_prev, -0.01, 1.0, 0.99, train).items():
            out[f"{k}_train{train}"] = v
    np.savez_compressed(os.path.join(HERE, "concrete.npz"), log_odds=lo, u=u, stop_prev=stop_prev,
                        loss_prev=loss_prev, digits_prev=dig_prev, prior=np.float32(-0.01), temperature=np.float32(1.0),
                        thr=np.float32(0.99), **out)

    # ---- full model, default hyper-parameters, weights/noise/images regenerated from seeds
    Bm = 8
    imgs, cnt = O.synthetic_canvases(Bm, seed=0)
    res = {}
    for train in (True, False):
        m = O.AIROracle(annealing_schedules=O.DEFAULT_ANNEALING, train=train, seed=0)
        m.global_step = 2000
        noise = O.make_noise(0, 3, Bm)
        out = m.forward(imgs, cnt, noise)
        tag = "train" if train else "test"
        for k in ("loss", "accuracy", "rec_num_digits", "reconstruction_loss", "rec_scales", "rec_shifts", "z_pres",
                  "z_pres_kls", "scale_kls", "shift_kls", "vae_kls", "stop_masks", "loss_per_item"):
            res[f"{tag}_{k}"] = out[k].numpy()
        res[f"{tag}_executed_steps"] = np.int32(out["executed_steps"])
    m = O.AIROracle(annealing_schedules=O.DEFAULT_ANNEALING, train=True, seed=0)
    m.global_step = 2000
    out, grads = m.loss_and_grads(imgs, cnt, O.make_noise(0, 3, Bm))
    for k, g_ in grads.items():
        res["gradnorm_" + k] = np.float32(g_.norm().item())
    np.savez_compressed(os.path.join(HERE, "model_b8.npz"), **res)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()

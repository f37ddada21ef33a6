"""Parity AT the BASELINE.json configurations themselves (kernel selection changes with the batch: images per CTA,
tile / split-K / CTA-pair heuristics of the GEMMs), not only on the small fixtures of the other GPU tests:

  configs[2]  full train step, B = 4096, the FP32-grade tensor-core mode (3xTF32), against the oracle
  configs[4]  inference, B = 65536, 5 steps: digit counts / stop masks / per-step outputs of every 64th image against
              the oracle (images are independent, so the oracle runs the sample only)
  configs[1]  the ST backward kernels at B = 65536 against the C oracle on a strided sample, flags 0 and 3
  the model's batched-over-T ST launches (air_st_*_steps) at T = 3, B = 4096 against the C oracle

All through the C ABI; bars as everywhere: integers / masks / ST forward bit-exact, <= 1e-5 forward, <= 1e-4 gradients."""
import numpy as np
import pytest
import torch

import air_b200 as ab
from air_b200 import ops
from oracle import air_oracle as O
from oracle import c_oracle as C
from tests.parity_util import covered_fixture, cuda_noise, make_pair, relnorm

pytestmark = pytest.mark.gpu
DEV = "cuda"
K = ab._cabi


def test_train_step_b4096_tf32x3_against_oracle():
    """configs[2] (B = 4096, T = 3) in the mode bench.py reports as exact_fp32_mode: digit counts and stop masks
    bit-exact, loss / per-step outputs <= 1e-5, all 36 gradients <= 1e-4 norm-wise."""
    B = 4096
    imgs, cnt, params, noise = covered_fixture(B, seed=21)
    orc, m = make_pair(imgs, cnt, params, train=True, gemm_mode="tf32x3")
    out, grads = orc.loss_and_grads(imgs, cnt, noise)
    m.loss_and_grads(cuda_noise(noise))
    assert torch.equal(m.rec_num_digits.cpu(), out["rec_num_digits"]) and torch.equal(m.stop_masks.cpu(), out["stop_masks"])
    assert abs(m.loss.item() - out["loss"].item()) <= 1e-5 * abs(out["loss"].item())
    for name in ("rec_scales", "rec_shifts", "rec_windows", "rec_latents", "z_pres_probs", "vae_kls", "reconstruction"):
        assert relnorm(getattr(m, name), out[name]) < 1e-5, (name, relnorm(getattr(m, name), out[name]))
    bad = {k: relnorm(g, grads[k]) for k, g in m.store.named_grads().items() if relnorm(g, grads[k]) > 1e-4}
    assert not bad, bad


def _realistic_params(seed):
    p = O.init_params(seed=seed)
    p["scale/mean/output/weights"] *= 3.0
    p["shift/mean/output/weights"] *= 3.0
    p["scale/log_variance/output/biases"] -= 3.0
    p["shift/log_variance/output/biases"] -= 3.0
    p["z_pres/log_odds/output/biases"] += 1.0
    p["vae/gen_mean/biases"] -= 1.5
    return p


@pytest.mark.parametrize("gemm_mode", ["tf32x3"])
def test_inference_b65536_five_steps_strided_sample_against_oracle(gemm_mode):
    """configs[4]: B = 65536, max_steps = 5, train=False.  Every 64th image (1024 of them) through the oracle with the
    same noise rows: digit counts (0..5 occur), stop masks and executed steps exact, per-step outputs <= 1e-5."""
    B, T, stride = 65536, 5, 64
    imgs, cnt = ab.data.device_canvases(B, seed=5)
    params = _realistic_params(31)
    g = torch.Generator(device=DEV).manual_seed(77)
    noise = dict(scale=torch.randn(T, B, 1, device=DEV, generator=g), shift=torch.randn(T, B, 2, device=DEV, generator=g),
                 vae_latent=torch.randn(T, B, 50, device=DEV, generator=g), vae_like=torch.randn(T, B, 784, device=DEV, generator=g),
                 concrete_u=torch.rand(T, B, device=DEV, generator=g))
    ab.reset_variable_scopes()
    h = dict(O.DEFAULT_HYPER)
    h["max_steps"] = T
    m = ab.AIRModel(imgs, cnt, train=False, annealing_schedules=O.DEFAULT_ANNEALING, gemm_mode=gemm_mode, **h)
    m.store.load_named({k: v.cuda() for k, v in params.items()})
    m.store.global_step = 0
    m.run(noise)
    sub = torch.arange(0, B, stride)
    orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING, train=False,
                      max_steps=T)
    orc.global_step = 0
    out = orc.forward(imgs[sub.to(DEV)].cpu(), cnt[sub.to(DEV)].cpu(), {k: v[:, sub.to(DEV)].cpu() for k, v in noise.items()})
    digits = m.rec_num_digits[sub.to(DEV)].cpu()
    assert torch.equal(digits, out["rec_num_digits"]) and torch.equal(m.stop_masks[sub.to(DEV)].cpu(), out["stop_masks"])
    assert len(torch.unique(digits)) >= 4, "the fixture should exercise several digit counts"
    for name in ("rec_scales", "rec_shifts", "rec_st_back", "rec_windows", "rec_latents", "z_pres_probs", "z_pres_kls", "scale_kls",
                 "shift_kls", "vae_kls"):
        got = getattr(m, name)[sub.to(DEV)].cpu()
        assert relnorm(got, out[name]) < 1e-5, (name, relnorm(got, out[name]))
    assert m.z_pres.unique().tolist() in ([0.0, 1.0], [0.0], [1.0])     # rounded at test time, whole batch
    del m
    torch.cuda.empty_cache()


def _poses(B, gen):
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)
    s = r(B) * 0.6 + 0.3
    xy = r(B, 2) - 0.5
    th = torch.zeros(B, 6, device=DEV)
    th[:, 0] = s; th[:, 4] = s; th[:, 2] = xy[:, 0]; th[:, 5] = xy[:, 1]
    thi = torch.zeros(B, 6, device=DEV)
    thi[:, 0] = 1 / s; thi[:, 4] = 1 / s; thi[:, 2] = -xy[:, 0] / s; thi[:, 5] = -xy[:, 1] / s
    return th, thi


def _oracle_wb_bwd(win, thi, z, stop, g, sigmoid_window):
    """C-oracle gradients of out = canvas + (stop < thr ? z * ST(win, thi) : 0) for the rows given (numpy)."""
    n = len(z)
    live = (stop < 0.99).astype(np.float32)
    up = (g * (z * live)[:, None, None]).astype(np.float32)
    dU, dth = C.st_backward(win.reshape(n, 28, 28, 1), thi.reshape(n, 2, 3), up.reshape(n, 50, 50, 1))
    dU = dU.reshape(n, 28, 28)
    if sigmoid_window:
        dU = dU * win * (1 - win)
    plain = C.st_forward(win.reshape(n, 28, 28, 1), thi.reshape(n, 2, 3), (50, 50)).reshape(n, 50, 50)
    dz = (g.astype(np.float64) * plain).reshape(n, -1).sum(1) * live
    return dU, dth.reshape(n, 6), dz


@pytest.mark.parametrize("flags", [0, 3])
def test_st_backward_kernels_b65536_strided_oracle_sample(flags):
    """configs[1] at its largest batch: air_st_writeback_canvas_bwd (flags 0: all six d theta^-1 entries; flags 3: the
    model's call, SigmoidGrad fused + axis-aligned theta) and air_st_backward (crop, d theta) against the C oracle on every
    512th row; stopped rows give exact zeros."""
    B, stride = 65536, 512
    gen = torch.Generator(device=DEV).manual_seed(3 + flags)
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)
    th, thi = _poses(B, gen)
    win, z = r(B, 28, 28) * 0.98 + 0.01, r(B)
    stop = (r(B) > 0.7).float() * 1.5
    g = torch.randn(B, 50, 50, device=DEV, generator=gen)
    dw, dt, dz = torch.empty_like(win), torch.empty(B, 6, device=DEV), torch.empty(B, device=DEV)
    L = K.lib()
    K.check(L.air_st_writeback_canvas_bwd(K.ptr(win), K.ptr(thi), K.ptr(z), K.ptr(stop), 0.99, K.ptr(g), K.ptr(dw), K.ptr(dt),
                                          K.ptr(dz), flags, B, 28, 28, 50, 50, K.stream()), "wb bwd")
    sub = torch.arange(0, B, stride, device=DEV)
    n = lambda t: t[sub].cpu().numpy()
    wU, wth, wz = _oracle_wb_bwd(n(win), n(thi), n(z), n(stop), n(g), bool(flags & 1))
    live = n(stop) < 0.99
    assert live.sum() > 50 and (~live).sum() > 20
    assert relnorm(n(dw), wU) < 1e-4 and relnorm(n(dz), wz) < 1e-4
    cols = [0, 2, 4, 5] if flags & 2 else list(range(6))
    assert relnorm(n(dt)[:, cols], wth[:, cols]) < 1e-4
    assert not n(dw)[~live].any() and not n(dt)[~live].any() and not n(dz)[~live].any()
    # crop backward (d theta only: the canvas is data)
    U, gw = r(B, 50, 50, 1), torch.randn(B, 28, 28, 1, device=DEV, generator=gen)
    dth = torch.empty(B, 6, device=DEV)
    K.check(L.air_st_backward(K.ptr(U), K.ptr(th), K.ptr(gw), None, K.ptr(dth), B, 50, 50, 1, 28, 28, K.stream()), "crop bwd")
    _, want = C.st_backward(n(U), n(th).reshape(-1, 2, 3), n(gw), need_dU=False)
    assert relnorm(n(dth), want.reshape(-1, 6)) < 1e-4


def test_steps_batched_st_launches_b4096_against_oracle():
    """The launches the model makes at configs[2] (T = 3, B = 4096): air_st_forward_steps / air_st_backward_steps (T crops of
    one canvas), air_st_writeback_canvas_fwd_steps (all T write-backs in one pass) and ..._bwd_steps.  Forward results
    bit-exact with the C oracle on ALL rows; backward against the oracle on every 16th row."""
    T, B, stride = 3, 4096, 16
    gen = torch.Generator(device=DEV).manual_seed(9)
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)
    U = (r(B, 2500) > 0.85).float() * r(B, 2500)
    th, thi = zip(*[_poses(B, gen) for _ in range(T)])
    th, thi = torch.stack(th), torch.stack(thi)
    win = torch.empty(T, B, 784, device=DEV)
    ops.st_forward_steps(U, th, win, 50, 50, 1, 28, 28)
    Un = U.cpu().numpy().reshape(B, 50, 50, 1)
    for t in range(T):
        assert np.array_equal(win[t].cpu().numpy().reshape(B, 28, 28, 1), C.st_forward(Un, th[t].cpu().numpy().reshape(B, 2, 3), (28, 28)))
    # fused write-backs, all T in one pass, against T oracle canvas updates
    recon = r(T, B, 784) * 0.98 + 0.01
    fields = torch.zeros(T, K.NF, B, device=DEV)
    fields[:, K.F_Z] = r(T, B)
    fields[:, K.F_STOP_NEW] = (r(T, B) > 0.7).float() * 1.5
    canvas = torch.empty(B, 2500, device=DEV)
    ops.writeback_canvas_fwd_steps(recon, thi, fields[0, K.F_Z], fields[0, K.F_STOP_NEW], K.NF * B, 0.99, None, canvas, 28, 28, 50, 50)
    want = np.zeros((B, 2500), np.float32)
    for t in range(T):
        wr = C.st_forward(recon[t].cpu().numpy().reshape(B, 28, 28, 1), thi[t].cpu().numpy().reshape(B, 2, 3), (50, 50)).reshape(B, 2500)
        want = C.canvas_update(want, wr, fields[t, K.F_Z].cpu().numpy(), fields[t, K.F_STOP_NEW].cpu().numpy(), 0.99)
    assert np.array_equal(canvas.cpu().numpy(), want)
    # backward of both, strided oracle sample
    dwin = torch.randn(T, B, 784, device=DEV, generator=gen)
    dth = torch.empty(T, B, 6, device=DEV)
    ops.st_backward_steps(U, th, dwin, dth, 50, 50, 1, 28, 28)
    dcanvas = torch.randn(B, 2500, device=DEV, generator=gen)
    dgen, dthi, dz = torch.empty(T, B, 784, device=DEV), torch.empty(T, B, 6, device=DEV), torch.empty(T, B, device=DEV)
    ops.writeback_canvas_bwd_steps(recon, thi, fields[0, K.F_Z], fields[0, K.F_STOP_NEW], K.NF * B, 0.99, dcanvas, dgen, dthi, dz,
                                   28, 28, 50, 50, window_is_sigmoid=True, axis_aligned_theta=True)
    sub = torch.arange(0, B, stride, device=DEV)
    n = lambda t: t[sub].cpu().numpy()
    for t in range(T):
        _, wth = C.st_backward(n(U).reshape(-1, 50, 50, 1), n(th[t]).reshape(-1, 2, 3), n(dwin[t]).reshape(-1, 28, 28, 1), need_dU=False)
        assert relnorm(n(dth[t]), wth.reshape(-1, 6)) < 1e-4
        wU, wti, wz = _oracle_wb_bwd(n(recon[t]).reshape(-1, 28, 28), n(thi[t]), n(fields[t, K.F_Z]), n(fields[t, K.F_STOP_NEW]),
                                     n(dcanvas).reshape(-1, 50, 50), True)
        assert relnorm(n(dgen[t]).reshape(-1, 28, 28), wU) < 1e-4 and relnorm(n(dz[t]), wz) < 1e-4
        assert relnorm(n(dthi[t])[:, [0, 2, 4, 5]], wti[:, [0, 2, 4, 5]]) < 1e-4

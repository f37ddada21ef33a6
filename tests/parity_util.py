"""Shared helpers for the parity tests and __graft_entry__.smoke() (test infrastructure)."""
import numpy as np
import torch

from oracle import air_oracle as O


def relnorm(a, b):
    a = a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if torch.is_tensor(b) else np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def covered_fixture(B, seed=0, T=3, cnn=False):
    """Well-conditioned ("covered", SURVEY hard part 2) fixture.  Every canvas pixel must sample
    INSIDE the window, otherwise the write-back leaves +-1e-9 rounding residues on it and
    max(canvas, 0) passes or blocks that pixel's gradient depending on the residue's sign (which
    flips with 1-ulp changes of theta).  In range <=> -1 <= (c - x)/s < 1.00007 for c in [-1, 1],
    so the fixture uses s >= 0.99999 (scale bias +12) and x = y = -3e-5 +- 1e-6 (shift mean bias,
    log-variance -30), all steps live.  The per-step reconstruction is < 0.1 so the 3-step canvas
    stays well inside (0, 1) (near 1 the BCE gradient 1/(1-r) is itself ill-conditioned), and
    digits are kept away from the border."""
    imgs, cnt = O.synthetic_canvases(B, seed=seed)
    im = imgs.reshape(B, 50, 50).clone()
    im[:, :7] = 0; im[:, -7:] = 0; im[:, :, :7] = 0; im[:, :, -7:] = 0
    params = O.init_params(seed=seed, cnn=cnn)
    params["scale/mean/output/biases"] += 12.0
    params["scale/mean/output/weights"] *= 0.2
    params["scale/log_variance/output/biases"] -= 8.0
    params["shift/mean/output/weights"] *= 0.0
    params["shift/mean/output/biases"] -= 3e-5
    params["shift/log_variance/output/weights"] *= 0.1
    params["shift/log_variance/output/biases"] -= 30.0
    params["z_pres/log_odds/output/biases"] += 5.0
    params["vae/gen_mean/weights"] *= 0.5
    params["vae/gen_mean/biases"] -= 3.5
    return im.reshape(B, -1).contiguous(), cnt, params, O.make_noise(seed, T, B)


def realistic_fixture(B, seed=0, T=3):
    """Poses like a trained model's (s ~ 0.35..0.75, shifts ~ +-0.3, mostly-live steps): windows
    smaller than the canvas, so most canvas pixels are out of range (residues).  Used with a
    teacher-forced dcanvas to check the backward schedule independently of the clip's sign noise."""
    imgs, cnt = O.synthetic_canvases(B, seed=seed)
    params = O.init_params(seed=seed)
    params["scale/mean/output/weights"] *= 3.0
    params["shift/mean/output/weights"] *= 3.0
    params["scale/log_variance/output/biases"] -= 3.0
    params["shift/log_variance/output/biases"] -= 3.0
    params["z_pres/log_odds/output/biases"] += 2.0
    params["vae/gen_mean/biases"] -= 1.5
    return imgs, cnt, params, O.make_noise(seed, T, B)


def default_fixture(B, seed=0, T=3):
    imgs, cnt = O.synthetic_canvases(B, seed=seed)
    return imgs, cnt, O.init_params(seed=seed), O.make_noise(seed, T, B)


def make_pair(imgs, cnt, params, train=True, global_step=2000, gemm_mode="fp32", reference_rounding=True, **hyper):
    """(oracle, cuda model) with identical parameters / hyper-parameters / global step."""
    import air_b200 as ab
    h = dict(O.DEFAULT_HYPER)
    h.update(hyper)
    orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING,
                      train=train, **hyper)
    orc.global_step = global_step
    ab.reset_variable_scopes()
    m = ab.AIRModel(imgs.cuda(), cnt.cuda(), train=train, annealing_schedules=O.DEFAULT_ANNEALING, gemm_mode=gemm_mode,
                    reference_rounding=reference_rounding, **{k: v for k, v in h.items()})
    m.store.load_named({k: v.cuda() for k, v in params.items()})
    m.store.global_step = global_step
    return orc, m


def cuda_noise(noise):
    return {k: v.cuda() for k, v in noise.items()}


def smoke_model():
    """One tiny train step of the full model on cuda:0 in each GEMM mode, checked against the oracle: the two FP32-grade
    modes (SIMT FFMA, 3xTF32 on tcgen05) at the parity bars, plain TF32 (the throughput mode, also tcgen05) at its own."""
    imgs, cnt, params, noise = covered_fixture(256, seed=0)    # 256 rows: the CTA-pair tensor-core kernels engage too
    orc = None
    for mode, loss_bar, grad_bar in (("fp32", 1e-5, 1e-4), ("tf32x3", 1e-5, 1e-4), ("tf32", 5e-3, 5e-2)):
        orc, m = make_pair(imgs, cnt, params, train=True, gemm_mode=mode)
        if mode == "fp32":
            out, grads = orc.loss_and_grads(imgs, cnt, noise)
        m.loss_and_grads(cuda_noise(noise))
        if mode != "tf32":
            assert torch.equal(m.rec_num_digits.cpu(), out["rec_num_digits"]), f"{mode}: digit counts differ from the oracle"
        rel = abs(m.loss.item() - out["loss"].item()) / abs(out["loss"].item())
        assert rel < loss_bar, f"{mode}: loss differs from the oracle by {rel:.2e}"
        worst = max(relnorm(g, grads[k]) for k, g in m.store.named_grads().items())
        assert worst < grad_bar, f"{mode}: gradient differs from the oracle by {worst:.2e}"
        m._apply_gradients()
        m.train_step()          # one step with noise generated on the device (counter-based RNG, in-epilogue likelihood noise)
        torch.cuda.synchronize()
        print(f"smoke model OK [{mode}]: loss rel err {rel:.2e}, worst grad rel err {worst:.2e}")

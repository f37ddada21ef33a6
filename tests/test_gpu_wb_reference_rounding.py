"""The write-back backward WITH THE REFERENCE'S ROUNDING (AIR_WB_REFERENCE_ROUNDING, csrc/st.cu st_wb_bwd_ref).

A canvas pixel outside the attention window meets one clipped border pixel of the window twice, with the weights (v - i)
and (i - v) (transformer.py:84-115).  On paper its gradient contributions are zero; the reference's fp32 autodiff
multiplies the upstream gradient into each corner term separately, and under the BCE loss the upstream gradient of a lit,
not yet reconstructed pixel is -x / (0 + 1e-9), so the reference sums ~1e10 products that cancel to ~1e3 rounding
residues -- and its training depends on them (test_training_needs_the_reference_rounding below).

What can be pinned numerically and what cannot:
  * per canvas pixel the arithmetic is the reference graph's own (gradients/AddN_10, AddN_11 of model/air-model.meta):
    dz and dtheta_inv are sums over pixels of per-pixel values restated here in numpy fp32, op for op -> equal up to the
    order of the final sum (tolerance: 1e-5 of the sum of absolute terms);
  * dwindow accumulates ~1e10 terms per border pixel whose ORDER the reference leaves open (UnsortedSegmentSum:
    sequential on a CPU, atomics on a GPU).  Pinned here: interior window pixels (fed by in-range canvas pixels only) equal
    the analytic kernel; border pixels differ from it by no more than the rounding of the terms they sum; the result is
    deterministic; with benign upstream gradients everything equals the C oracle at the usual 1e-4."""
import numpy as np
import pytest
import torch

import air_b200 as ab
from air_b200 import ops
from oracle import c_oracle as C
from tests.parity_util import relnorm

pytestmark = pytest.mark.gpu
DEV = "cuda"
K = ab._cabi
F = np.float32


def _poses(B, rng, s_lo=0.3, s_hi=0.8, shift=0.5):
    s = rng.uniform(s_lo, s_hi, B).astype(F)
    x = rng.uniform(-shift, shift, B).astype(F)
    y = rng.uniform(-shift, shift, B).astype(F)
    z = np.zeros(B, F)
    return np.stack([F(1) / s, z, -x / s, z, F(1) / s, -y / s], 1).astype(F)   # three separate divisions (air_model.py:351-360)


def _run(win, thi, z, stop, g, flags):
    B = len(z)
    win, thi, z, stop, g = (torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in (win, thi, z, stop, g))  # (kept alive)
    dw = torch.full((B, 28, 28), 7.0, device=DEV); dt = torch.full((B, 6), 7.0, device=DEV); dz = torch.full((B,), 7.0, device=DEV)
    K.check(K.lib().air_st_writeback_canvas_bwd(K.ptr(win), K.ptr(thi), K.ptr(z), K.ptr(stop), 0.99, K.ptr(g),
                                                K.ptr(dw), K.ptr(dt), K.ptr(dz), flags, B, 28, 28, 50, 50, K.stream()), "wb bwd")
    return dw.cpu().numpy(), dt.cpu().numpy(), dz.cpu().numpy()


def _tables(diag, trans, n_out, n_in):
    """Clipped corners and 1-D weights of one axis, fp32 op for op (transformer.py:75-115, 127-130, 159)."""
    k = np.arange(n_out, dtype=F)
    gk = (F(-1.0) + (F(2.0) / F(n_out - 1)) * k).astype(F)
    v = (diag[:, None] * gk[None, :]).astype(F) + trans[:, None]
    px = ((v + F(1.0)).astype(F) * (F(n_in) - F(1.001))).astype(F) * F(0.5)
    fl = np.floor(px).astype(np.int64)
    i0, i1 = np.clip(fl, 0, n_in - 1), np.clip(fl + 1, 0, n_in - 1)
    return gk, i0, i1, (i1.astype(F) - px).astype(F), (px - i0.astype(F)).astype(F)


def _graph_per_pixel(win, thi, z, d):
    """dz, dtheta_inv as the reference graph forms them PER PIXEL (fp32, every op rounded), summed over pixels in fp64;
    also the sums of absolute terms (the scale the summation order can move the fp32 result by)."""
    B = len(z)
    xt, j0, j1, cw1, cw0 = _tables(thi[:, 0], thi[:, 2], 50, 28)
    yt, i0, i1, rw1, rw0 = _tables(thi[:, 4], thi[:, 5], 50, 28)
    bi = np.arange(B)[:, None, None]
    Ia, Ib = win[bi, i0[:, :, None], j0[:, None, :]], win[bi, i1[:, :, None], j0[:, None, :]]
    Ic, Id = win[bi, i0[:, :, None], j1[:, None, :]], win[bi, i1[:, :, None], j1[:, None, :]]
    X1, X0, Y1, Y0 = cw1[:, None, :], cw0[:, None, :], rw1[:, :, None], rw0[:, :, None]
    wa, wb, wc, wd = (X1 * Y1).astype(F), (X1 * Y0).astype(F), (X0 * Y1).astype(F), (X0 * Y0).astype(F)
    sample = ((wa * Ia + wb * Ib).astype(F) + wc * Ic).astype(F) + wd * Id
    g = (d * z[:, None, None]).astype(F)
    da, db, dc, dd = g * Ia, g * Ib, g * Ic, g * Id
    dx = (((-(da * Y1)) + (-(db * Y0))).astype(F) + dc * Y1).astype(F) + dd * Y0
    dy = (((-(da * X1)) + db * X1).astype(F) + (-(dc * X0))).astype(F) + dd * X0
    f8 = np.float64
    sx, sy = f8(F(0.5) * (F(28) - F(1.001))), f8(F(0.5) * (F(28) - F(1.001)))
    dx8, dy8 = dx.astype(f8), dy.astype(f8)
    XT, YT = xt.astype(f8)[None, None, :], yt.astype(f8)[None, :, None]
    dth = np.stack([(dx8 * XT).sum((1, 2)) * sx, (dx8 * YT).sum((1, 2)) * sx, dx8.sum((1, 2)) * sx,
                    (dy8 * XT).sum((1, 2)) * sy, (dy8 * YT).sum((1, 2)) * sy, dy8.sum((1, 2)) * sy], 1)
    scale_th = np.stack([np.abs(dx8).sum((1, 2)) * sx] * 3 + [np.abs(dy8).sum((1, 2)) * sy] * 3, 1)
    dzs = (d.astype(f8) * sample.astype(f8)).sum((1, 2))
    scale_z = np.abs(d.astype(f8) * sample.astype(f8)).sum((1, 2))
    # sum of the absolute corner terms every window pixel receives (bound for dwindow's cancellation residue)
    terms = np.zeros((B, 28, 28), f8)
    for w_, ii, jj in ((wa, i0, j0), (wb, i1, j0), (wc, i0, j1), (wd, i1, j1)):
        np.add.at(terms, (bi, ii[:, :, None], jj[:, None, :]), np.abs(g.astype(f8) * w_.astype(f8)))
    return dth, scale_th, dzs, scale_z, sample, terms


def _spiky_case(B, seed):
    """Upstream gradient of the BCE loss on canvases the window does not cover: -x / (clip(canvas) + 1e-9) on lit pixels
    whose canvas value passes the clip (>= 0), the canvas being the write-back itself (residues included)."""
    rng = np.random.RandomState(seed)
    thi = _poses(B, rng)
    win = (1.0 / (1.0 + np.exp(-rng.randn(B, 28, 28) * 0.5))).astype(F)
    z = rng.uniform(0.3, 1.0, B).astype(F)
    stop = np.where(rng.rand(B) < 0.2, 1.5, 0.0).astype(F)
    canvas = (C.st_forward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), (50, 50)).reshape(B, 50, 50) * z[:, None, None]).astype(F)
    lit = rng.rand(B, 50, 50) < 0.12
    x = np.where(lit, rng.uniform(0.3, 1.0, (B, 50, 50)), 0.0).astype(F)
    rec = np.clip(canvas, 0.0, 1.0)
    d = (-(x / (rec + F(1e-9))) + (F(1) - x) / (F(1) - rec + F(1e-9))).astype(F)
    d = np.where((canvas >= 0) & (canvas <= 1), d, 0).astype(F) / F(64)
    return win, thi, z, stop, d


def test_smooth_upstream_gradient_equals_the_oracle():
    """Benign upstream gradients (no 1e9 spikes): every output of the reference-rounding kernel -- all six dtheta_inv
    entries, dwindow with and without the fused SigmoidGrad, dz -- equals the C oracle at the usual bars."""
    B = 384
    rng = np.random.RandomState(3)
    thi = _poses(B, rng)
    win = rng.uniform(0.01, 0.99, (B, 28, 28)).astype(F)
    z = rng.uniform(0.1, 1.0, B).astype(F)
    stop = np.where(rng.rand(B) < 0.25, 1.5, 0.0).astype(F)
    g = rng.randn(B, 50, 50).astype(F)
    live = stop < 0.99
    up = (g * (z * live)[:, None, None]).astype(F)
    wU, wth = C.st_backward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), up.reshape(B, 50, 50, 1))
    wU, wth = wU.reshape(B, 28, 28), wth.reshape(B, 6)
    plain = C.st_forward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), (50, 50)).reshape(B, 50, 50)
    wz = (g.astype(np.float64) * plain).reshape(B, -1).sum(1) * live
    for flags in (4, 5, 6, 7):
        dw, dt, dz = _run(win, thi, z, stop, g, flags)
        want = wU * win * (1 - win) if flags & 1 else wU
        assert relnorm(dw, want) < 1e-4 and relnorm(dz, wz) < 1e-4 and relnorm(dt, wth) < 1e-4, flags
        assert not dw[~live].any() and not dt[~live].any() and not dz[~live].any()


def test_spiky_upstream_gradient_per_pixel_terms_are_the_graphs():
    B = 96
    win, thi, z, stop, d = _spiky_case(B, 11)
    live = stop < 0.99
    dth, scale_th, dzs, scale_z, sample, terms = _graph_per_pixel(win, thi, z, d)
    assert np.abs(d).max() > 1e6                                    # the spikes are there
    assert ((np.abs(sample) < 1e-5) & (sample != 0)).mean() > 0.03  # ... and so are the forward residues
    dw, dt, dz = _run(win, thi, z, stop, d, 4)
    dw2, dt2, dz2 = _run(win, thi, z, stop, d, 4)
    assert np.array_equal(dw, dw2) and np.array_equal(dt, dt2) and np.array_equal(dz, dz2)   # deterministic
    L = live
    assert (np.abs(dt[L] - dth[L]) <= 2e-5 * scale_th[L] + 1e-6).all(), np.abs(dt[L] - dth[L]).max()
    assert (np.abs(dz[L] - dzs[L]) <= 2e-5 * scale_z[L] + 1e-6).all()
    assert not dw[~L].any() and not dt[~L].any() and not dz[~L].any()
    # the un-cancelled terms are really summed: the analytic kernel (exact cancellation) gives another dtheta_inv / dz
    cw, ct, cz = _run(win, thi, z, stop, d, 2)
    assert np.abs(dt[L] - ct[L])[:, [0, 2, 4, 5]].max() > 1.0 and np.abs(dz[L] - cz[L]).max() > 1e-3
    # dwindow: interior pixels are fed by in-range canvas pixels only -> the analytic result; border pixels carry the
    # cancellation residue of the terms they sum, bounded by a few ulps of the sum of absolute terms
    inner = np.zeros((28, 28), bool); inner[1:-1, 1:-1] = True
    assert relnorm(dw[L][:, inner], cw[L][:, inner]) < 1e-5
    resid = np.abs(dw.astype(np.float64) - cw)[L]
    assert (resid <= 512 * 2.0 ** -24 * terms[L] + 1e-4 * np.abs(cw[L]) + 1e-6).all()
    assert resid[:, ~inner].max() > 10.0                            # residues of the reference's magnitude are present
    # ... the magnitude the reference graph's own accumulation order leaves (oracle_st_backward is bit-exact with the
    # graph's UnsortedSegmentSum order): same order of magnitude, image by image in the median
    up = (d * z[:, None, None]).astype(F)
    oU, _ = C.st_backward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), up.reshape(B, 50, 50, 1))
    o_res = np.abs(oU.reshape(B, 28, 28).astype(np.float64) - cw)[L][:, ~inner]
    ours, theirs = np.sqrt((resid[:, ~inner] ** 2).mean(1)), np.sqrt((o_res ** 2).mean(1))
    ok = theirs > 1.0
    ratio = np.median(ours[ok] / theirs[ok])
    print("border residue rms, ours / graph order: median ratio", ratio, " (ours", np.median(ours[ok]), ", graph order", np.median(theirs[ok]), ")")
    assert ok.sum() > 20 and 0.03 < ratio < 30.0


def test_rotated_theta_takes_the_per_pixel_path():
    """A theta_inv with shear / rotation / a mirrored window cannot use the separable scans: per-pixel path with the same
    per-pixel arithmetic (dwindow through shared-memory atomics)."""
    B = 32
    rng = np.random.RandomState(5)
    thi = _poses(B, rng)
    thi[::3, 1] = rng.uniform(-0.3, 0.3, len(thi[::3])).astype(F)
    thi[1::3, 0] *= -1
    win = rng.uniform(0.01, 0.99, (B, 28, 28)).astype(F)
    z = rng.uniform(0.1, 1.0, B).astype(F)
    stop = np.zeros(B, F)
    g = rng.randn(B, 50, 50).astype(F)
    up = (g * z[:, None, None]).astype(F)
    wU, wth = C.st_backward(win.reshape(B, 28, 28, 1), thi.reshape(B, 2, 3), up.reshape(B, 50, 50, 1))
    dw, dt, dz = _run(win, thi, z, stop, g, 4)
    # (the un-cancelled corner terms of the clipped pixels carry weights up to ~1e3: their rounding is the 1e-4)
    assert relnorm(dw, wU.reshape(B, 28, 28)) < 3e-4 and relnorm(dt, wth.reshape(B, 6)) < 1e-4


def test_steps_entry_point_bit_identical_to_single_steps():
    T, B = 3, 40
    rng = np.random.RandomState(9)
    win, thi, z, stop, d = _spiky_case(T * B, 21)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    NF = 4
    fields = torch.zeros(T, NF, B, device=DEV)
    fields[:, 0] = cu(z.reshape(T, B)); fields[:, 1] = cu(stop.reshape(T, B))
    recon, ti, dc = cu(win.reshape(T, B, 784)), cu(thi.reshape(T, B, 6)), cu(d[:B].reshape(B, 2500))
    dw, dt, dz = torch.empty(T, B, 784, device=DEV), torch.empty(T, B, 6, device=DEV), torch.empty(T, B, device=DEV)
    ops.writeback_canvas_bwd_steps(recon, ti, fields[0, 0], fields[0, 1], NF * B, 0.99, dc, dw, dt, dz, 28, 28, 50, 50,
                                   window_is_sigmoid=True, axis_aligned_theta=True, reference_rounding=True)
    for t in range(T):
        w1, t1, z1 = _run(win.reshape(T, B, 28, 28)[t], thi.reshape(T, B, 6)[t], z.reshape(T, B)[t], stop.reshape(T, B)[t],
                          d[:B], 5)
        assert np.array_equal(dw[t].cpu().numpy().reshape(B, 28, 28), w1) and np.array_equal(dt[t].cpu().numpy(), t1)
        assert np.array_equal(dz[t].cpu().numpy(), z1)


def test_training_needs_the_reference_rounding():
    """The reference's training configuration (training.py:100-122) on synthetic canvases, 6000 optimisation steps through
    the CUDA path: with the reference's rounding the reconstruction is learnt (loss < 700 from ~2000, as the CPU oracle's
    profiles/r1_oracle_convergence.log), with the exactly-cancelled gradient the loss has not moved."""
    data = ab.data
    train, cnt = data.device_canvases(20000, seed=0)
    final = {}
    for ref in (True, False):
        ab.reset_variable_scopes()
        m = ab.AIRModel(train[:64].clone(), cnt[:64].clone(), train=True, annealing_schedules=data.TRAINING_ANNEALING,
                        gemm_mode="tf32x3", seed=0, reference_rounding=ref, **data.TRAINING_HYPER)
        m.capture()
        import gc
        assert gc.isenabled()     # (capture() keeps the collector off only while the graph is being recorded)
        g = torch.Generator(device=DEV).manual_seed(1)
        tot = torch.zeros((), device=DEV)
        for it in range(6000):
            idx = torch.randint(0, 20000, (64,), generator=g, device=DEV)
            m.feed(train[idx], cnt[idx])
            m.train_step()
            if it >= 5500:
                tot += m.loss.reshape(())
        final[ref] = tot.item() / 500
    assert final[True] < 700.0 and final[False] > 1500.0, final

"""GPU parity tests (through the C ABI) for the Spatial Transformer, the fused write-back +
canvas kernel and the Concrete/ACT step, against the CPU oracle and the golden vectors.

Bars: forward ST / canvas = BIT-EXACT; Concrete ints/masks exact, floats <= 1e-5 rel;
gradients <= 1e-4 norm-wise relative."""
import os

import numpy as np
import pytest
import torch

import air_b200 as ab
from air_b200 import ops
from oracle import air_oracle as O
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def relnorm(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def air_thetas(rng, B, lo=0.3, hi=0.9):
    s = rng.uniform(lo, hi, B).astype(np.float32)
    x = rng.uniform(-0.5, 0.5, B).astype(np.float32)
    y = rng.uniform(-0.5, 0.5, B).astype(np.float32)
    th = np.zeros((B, 2, 3), np.float32)
    th[:, 0, 0] = s; th[:, 1, 1] = s; th[:, 0, 2] = x; th[:, 1, 2] = y
    thi = np.zeros((B, 2, 3), np.float32)
    one = np.float32(1)
    thi[:, 0, 0] = one / s; thi[:, 1, 1] = one / s; thi[:, 0, 2] = -x / s; thi[:, 1, 2] = -y / s
    return th, thi


def test_golden_forward_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "st.npz"))
    assert np.array_equal(ab.transformer(cu(g["U"]), cu(g["theta"]), (28, 28)).cpu().numpy(), g["crop"])
    assert np.array_equal(ab.transformer(cu(g["window"]), cu(g["theta_inv"]), (50, 50)).cpu().numpy(), g["back"])
    assert np.array_equal(ab.transformer(cu(g["Ug"]), cu(g["thg"]), (7, 9)).cpu().numpy(), g["outg"])
    out = ab.writeback_canvas(cu(g["window"][..., 0]), cu(g["theta_inv"]), cu(g["z"]), cu(g["stop"]),
                              cu(g["canvas"]).reshape(-1, 50, 50), 0.99)
    assert np.array_equal(out.cpu().numpy().reshape(-1, 2500), g["canvas_out"])


@pytest.mark.parametrize("B", [1, 3, 4, 5, 257, 4096])
def test_crop_forward_bit_exact(B):
    rng = np.random.RandomState(B)
    U = rng.rand(B, 50, 50, 1).astype(np.float32)
    th, _ = air_thetas(rng, B, 0.05, 1.6)   # includes windows larger than the canvas (clipping)
    got = ab.transformer(cu(U), cu(th), (28, 28)).cpu().numpy()
    assert np.array_equal(got, C.st_forward(U, th, (28, 28)))


@pytest.mark.parametrize("B", [1, 2, 7, 1000])
def test_writeback_forward_bit_exact(B):
    rng = np.random.RandomState(100 + B)
    win = rng.rand(B, 28, 28, 1).astype(np.float32)
    _, thi = air_thetas(rng, B, 0.1, 1.2)
    got = ab.transformer(cu(win), cu(thi), (50, 50)).cpu().numpy()
    assert np.array_equal(got, C.st_forward(win, thi, (50, 50)))


def test_mixed_separable_and_rotated_thetas_bit_exact():
    rng = np.random.RandomState(7)
    B = 37
    U = rng.rand(B, 50, 50, 1).astype(np.float32)
    th, _ = air_thetas(rng, B)
    th[::3] = rng.uniform(-1.1, 1.1, th[::3].shape)       # every third image: shear/rotation
    th[1, 0, 1] = -0.0                                     # negative zero still counts as axis-aligned
    got = ab.transformer(cu(U), cu(th), (28, 28)).cpu().numpy()
    assert np.array_equal(got, C.st_forward(U, th, (28, 28)))


@pytest.mark.parametrize("shape", [(5, 11, 13, 3, 7, 9), (2, 64, 48, 1, 33, 17), (3, 9, 9, 1, 100, 100),
                                   (2, 16, 16, 4, 1, 1), (2, 300, 300, 1, 20, 20)])
def test_generic_shapes_bit_exact(shape):
    B, H, W, Cc, oh, ow = shape
    rng = np.random.RandomState(sum(shape))
    U = rng.rand(B, H, W, Cc).astype(np.float32)
    th = rng.uniform(-1.2, 1.2, (B, 2, 3)).astype(np.float32)
    got = ab.transformer(cu(U), cu(th), (oh, ow)).cpu().numpy()
    assert np.array_equal(got, C.st_forward(U, th, (oh, ow)))


def test_unaligned_view_uses_generic_path_bit_exact():
    rng = np.random.RandomState(3)
    buf = cu(rng.rand(1 + 4 * 2500).astype(np.float32))
    U = buf[1:].reshape(4, 50, 50, 1)                      # 4-byte aligned only
    th, _ = air_thetas(rng, 4)
    got = ab.transformer(U, cu(th), (28, 28)).cpu().numpy()
    assert np.array_equal(got, C.st_forward(U.cpu().numpy(), th, (28, 28)))


def test_batch_transformer_and_empty():
    rng = np.random.RandomState(5)
    U = rng.rand(3, 50, 50, 1).astype(np.float32)
    th = np.stack([air_thetas(rng, 3)[0] for _ in range(2)], 1).reshape(3, 2, 6)
    got = ab.batch_transformer(cu(U), cu(th), (28, 28)).cpu().numpy()
    want = O.batch_transformer(torch.from_numpy(U), torch.from_numpy(th), (28, 28)).numpy()
    assert np.array_equal(got, want)
    assert ab.transformer(cu(U[:0]), cu(th[:0, 0]), (28, 28)).shape == (0, 28, 28, 1)


@pytest.mark.parametrize("inplace", [False, True])
def test_fused_canvas_forward_bit_exact(inplace):
    rng = np.random.RandomState(11)
    B = 513
    win = rng.rand(B, 28, 28).astype(np.float32)
    _, thi = air_thetas(rng, B)
    z = rng.rand(B).astype(np.float32)
    stop = rng.choice(np.array([0.0, 0.5, 0.98999, 0.99, 1.7], np.float32), B)
    canvas = rng.rand(B, 2500).astype(np.float32)
    want = C.canvas_update(canvas, C.st_forward(win[..., None], thi, (50, 50)).reshape(B, 2500), z, stop, 0.99)
    cv = cu(canvas).reshape(B, 50, 50)
    if inplace:
        l = ab._cabi
        keep = (cu(win), cu(thi).reshape(B, 6), cu(z), cu(stop))   # keep the device buffers alive for the launch
        l.check(l.lib().air_st_writeback_canvas_fwd(l.ptr(keep[0]), l.ptr(keep[1]), l.ptr(keep[2]), l.ptr(keep[3]), 0.99,
                                                   l.ptr(cv), l.ptr(cv), B, 28, 28, 50, 50, l.stream()), "inplace")
        got = cv
    else:
        got = ab.writeback_canvas(cu(win), cu(thi), cu(z), cu(stop), cv, 0.99)
    assert np.array_equal(got.cpu().numpy().reshape(B, 2500), want)


def _oracle_st_grads(U, th, g, out_size):
    Ut, tt = torch.from_numpy(U).requires_grad_(True), torch.from_numpy(th).requires_grad_(True)
    O.transformer(Ut, tt, out_size).backward(torch.from_numpy(g))
    return Ut.grad.numpy(), tt.grad.numpy()


def test_crop_backward_dtheta():
    rng = np.random.RandomState(21)
    B = 64
    imgs, _ = O.synthetic_canvases(B, seed=4)
    U = imgs.numpy().reshape(B, 50, 50, 1)
    th, _ = air_thetas(rng, B)
    g = rng.randn(B, 28, 28, 1).astype(np.float32)
    Ut, tt = cu(U), cu(th).requires_grad_(True)
    ab.transformer(Ut, tt, (28, 28)).backward(cu(g))
    _, dth = _oracle_st_grads(U, th, g, (28, 28))
    assert relnorm(tt.grad.cpu().numpy(), dth) < 1e-4
    # the C oracle (same formulas, sequential sums) agrees as well
    assert relnorm(tt.grad.cpu().numpy(), C.st_backward(U, th, g, need_dU=False)[1]) < 1e-4


def test_writeback_backward_dU_dtheta_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "st.npz"))
    Ut, tt = cu(g["window"]).requires_grad_(True), cu(g["theta_inv"]).requires_grad_(True)
    ab.transformer(Ut, tt, (50, 50)).backward(cu(g["back_dout"]))
    assert relnorm(Ut.grad.cpu().numpy(), g["back_dU"]) < 1e-4
    assert relnorm(tt.grad.cpu().numpy(), g["back_dtheta"]) < 1e-4


def test_backward_is_closer_to_fp64_truth_than_oracle_fp32():
    """dU on random upstream gradients is cancellation-dominated at the window border
    (SURVEY hard part 2/3): show the kernel is at least as close to fp64 as the fp32 oracle."""
    rng = np.random.RandomState(33)
    B = 16
    win = rng.rand(B, 28, 28, 1).astype(np.float32)
    _, thi = air_thetas(rng, B)
    g = rng.randn(B, 50, 50, 1).astype(np.float32)
    Ud, td = torch.from_numpy(win).double().requires_grad_(True), torch.from_numpy(thi).double().requires_grad_(True)
    O.transformer(Ud, td, (50, 50)).backward(torch.from_numpy(g).double())
    dU32, dth32 = _oracle_st_grads(win, thi, g, (50, 50))
    Ut, tt = cu(win).requires_grad_(True), cu(thi).requires_grad_(True)
    ab.transformer(Ut, tt, (50, 50)).backward(cu(g))
    e_gpu, e_orc = relnorm(Ut.grad.cpu().numpy(), Ud.grad.numpy()), relnorm(dU32, Ud.grad.numpy())
    assert e_gpu < max(2 * e_orc, 1e-5), (e_gpu, e_orc)
    assert relnorm(tt.grad.cpu().numpy(), td.grad.numpy()) < 1e-4


def test_backward_deterministic():
    rng = np.random.RandomState(34)
    B = 64
    win, g = rng.rand(B, 28, 28, 1).astype(np.float32), rng.randn(B, 50, 50, 1).astype(np.float32)
    _, thi = air_thetas(rng, B)
    res = []
    for _ in range(3):
        Ut, tt = cu(win).requires_grad_(True), cu(thi).requires_grad_(True)
        ab.transformer(Ut, tt, (50, 50)).backward(cu(g))
        res.append((Ut.grad.cpu().numpy(), tt.grad.cpu().numpy()))
    for r in res[1:]:
        assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1])


def test_generic_backward_rotation_multichannel():
    rng = np.random.RandomState(35)
    U = rng.rand(4, 12, 10, 3).astype(np.float32)
    th = (np.array([[0.8, 0.2, 0.0, -0.2, 0.7, 0.1]], np.float32) + 0.05 * rng.randn(4, 6)).astype(np.float32)
    g = rng.randn(4, 6, 7, 3).astype(np.float32)
    Ut, tt = cu(U).requires_grad_(True), cu(th).requires_grad_(True)
    ab.transformer(Ut, tt, (6, 7)).backward(cu(g))
    dU, dth = _oracle_st_grads(U, th, g, (6, 7))
    assert relnorm(Ut.grad.cpu().numpy(), dU) < 1e-4 and relnorm(tt.grad.cpu().numpy(), dth) < 1e-4
    # rotated single-channel image goes through the staged kernel's shared-memory-atomic path
    U1, g1 = U[..., :1].copy(), g[..., :1].copy()
    U1 = np.ascontiguousarray(np.pad(U1, ((0, 0), (0, 0), (0, 2), (0, 0))))   # 12x12: H*W % 4 == 0
    Ut, tt = cu(U1).requires_grad_(True), cu(th).requires_grad_(True)
    ab.transformer(Ut, tt, (6, 8)).backward(cu(np.pad(g1, ((0, 0), (0, 0), (0, 1), (0, 0)))))
    dU, dth = _oracle_st_grads(U1, th, np.pad(g1, ((0, 0), (0, 0), (0, 1), (0, 0))), (6, 8))
    assert relnorm(Ut.grad.cpu().numpy(), dU) < 1e-4 and relnorm(tt.grad.cpu().numpy(), dth) < 1e-4


def test_fused_canvas_backward():
    rng = np.random.RandomState(41)
    B = 96
    win = rng.rand(B, 28, 28).astype(np.float32)
    _, thi = air_thetas(rng, B)
    z = rng.rand(B).astype(np.float32)
    stop = rng.choice(np.array([0.0, 0.5, 0.99, 1.7], np.float32), B)
    canvas = rng.rand(B, 50, 50).astype(np.float32)
    # smooth upstream gradient (a "covered" fixture: no 1e9 amplification)
    g = (rng.rand(B, 50, 50).astype(np.float32) - 0.3)
    wt, tt, zt, ct = (torch.from_numpy(a).requires_grad_(True) for a in (win, thi, z, canvas))
    wr = O.transformer(wt.unsqueeze(3), tt, (50, 50))[..., 0]
    live = torch.from_numpy(stop) < 0.99
    out = ct + torch.where(live[:, None, None], zt[:, None, None] * wr, torch.zeros_like(wr))
    out.backward(torch.from_numpy(g))
    wg, tg, zg, cg = (cu(a).requires_grad_(True) for a in (win, thi, z, canvas))
    ab.writeback_canvas(wg, tg, zg, cu(stop), cg, 0.99).backward(cu(g))
    assert relnorm(wg.grad.cpu().numpy(), wt.grad.numpy()) < 1e-4
    assert relnorm(tg.grad.cpu().numpy(), tt.grad.numpy()) < 1e-4
    assert relnorm(zg.grad.cpu().numpy(), zt.grad.numpy()) < 1e-4
    assert np.array_equal(cg.grad.cpu().numpy(), g)
    dead = ~live.numpy()
    assert not wg.grad.cpu().numpy()[dead].any() and not zg.grad.cpu().numpy()[dead].any()


def _axis_cases(rng):
    """theta_inv rows [a, 0, tx, 0, e, ty] exercising every branch of the warp-specialised kernel."""
    rows = []
    for _ in range(40):                                   # the model's regime: window 30-90 % of the canvas
        s = rng.uniform(0.3, 0.9); x, y = rng.uniform(-0.5, 0.5, 2)
        rows.append([1 / s, 0, -x / s, 0, 1 / s, -y / s])
    for s in (1.0, 1.3, 2.0, 3.0):                       # window covers the canvas: 50 in-range rows / columns (2 blocks)
        rows.append([1 / s, 0, 0.02, 0, 1 / s, -0.03])
    for s in (0.04, 0.08, 0.15):                         # tiny windows: a few columns, source index jumps by > 1
        rows.append([1 / s, 0, 0.3 / s, 0, 1 / s, -0.2 / s])
    for x, y in ((0.9, 0.0), (-0.95, 0.4), (0.0, 0.97), (0.8, -0.85)):   # partially outside the canvas
        rows.append([2.0, 0, -2.0 * x, 0, 2.0, -2.0 * y])
    rows.append([2.0, 0, 7.0, 0, 2.0, 0.0])               # entirely outside: every gradient is zero
    rows.append([-1.6, 0, 0.1, 0, 1.7, 0.05])             # mirrored columns
    rows.append([1.4, 0, -0.2, 0, -2.1, 0.1])             # mirrored rows
    rows.append([1.5, 0, 0.1, 0, 0.7, -0.1])              # anisotropic
    rows.append([1.5173, 0.2291, 0.1037, -0.1113, 1.4219, 0.0071])   # shear (no sample lands exactly on the hard edge): falls back to the atomics path (all six entries)
    return np.asarray(rows, np.float32)


@pytest.mark.parametrize("sig", [0, 1])
def test_fused_canvas_backward_axis_aligned_kernel(sig):
    """AIR_WB_AXIS_ALIGNED_THETA: the P/Q run-scan kernel (dU, dz, dtheta[0,2,4,5]) against the oracle's autograd
    evaluated in float64 -- the fp32 oracle itself carries the reference's cancellation noise from clipped pixels
    (weights of +-300 on the tiny-window rows), which is not what this test is about -- and, on the model-regime
    rows, against the fp32 oracle at the parity tolerance."""
    rng = np.random.RandomState(43)
    thi = _axis_cases(rng)
    B = thi.shape[0]
    win = rng.rand(B, 28, 28).astype(np.float32) * 0.9 + 0.05
    z = (rng.rand(B).astype(np.float32) * 0.9 + 0.1)
    stop = rng.choice(np.array([0.0, 0.5, 1.7], np.float32), B, p=[0.45, 0.45, 0.1])
    g = (rng.rand(B, 50, 50).astype(np.float32) - 0.3)
    pre = np.log(win / (1 - win)).astype(np.float32)       # pre-sigmoid input of the window (sig = 1)
    wwin = torch.sigmoid(torch.from_numpy(pre)).numpy() if sig else win

    def oracle_grads(dtype):
        if sig:
            pt = torch.from_numpy(pre).to(dtype).requires_grad_(True)
            # the kernel multiplies by w (1 - w) of the fp32 window it is given
            wt = torch.from_numpy(wwin).to(dtype) + (torch.sigmoid(pt) - torch.sigmoid(pt).detach())
        else:
            pt = torch.from_numpy(win).to(dtype).requires_grad_(True)
            wt = pt
        tt, zt = torch.from_numpy(thi).to(dtype).requires_grad_(True), torch.from_numpy(z).to(dtype).requires_grad_(True)
        wr = O.transformer(wt.unsqueeze(3), tt, (50, 50))[..., 0]
        live = torch.from_numpy(stop) < 0.99
        (torch.where(live[:, None, None], zt[:, None, None] * wr, torch.zeros_like(wr)) * torch.from_numpy(g).to(dtype)).sum().backward()
        return pt.grad.numpy(), tt.grad.numpy(), zt.grad.numpy()

    want_w, want_t, want_z = oracle_grads(torch.float64)
    f32_w, f32_t, f32_z = oracle_grads(torch.float32)
    L, c = ab._cabi.lib(), ab._cabi
    wi, ti, zi, si, gi = cu(wwin), cu(thi), cu(z), cu(stop), cu(g)
    res = []
    for flags in (sig, sig | 2):
        dw = torch.full((B, 28, 28), 7.0, device=DEV); dt = torch.full((B, 6), 7.0, device=DEV); dz = torch.full((B,), 7.0, device=DEV)
        c.check(L.air_st_writeback_canvas_bwd(c.ptr(wi), c.ptr(ti), c.ptr(zi), c.ptr(si), 0.99, c.ptr(gi), c.ptr(dw),
                                              c.ptr(dt), c.ptr(dz), flags, B, 28, 28, 50, 50, c.stream()), "bwd")
        res.append((dw.cpu().numpy().astype(np.float64), dt.cpu().numpy().astype(np.float64), dz.cpu().numpy().astype(np.float64)))
    shear = thi[:, 1] != 0
    diag = [0, 2, 4, 5]
    model_rows = np.arange(B) < 40
    for name, (dw, dt, dz) in zip(("staged", "axis"), res):
        # per image, so that one bad branch cannot hide in a batch norm
        ew = np.abs(dw - want_w).reshape(B, -1).max(1) / (np.abs(want_w).reshape(B, -1).max(1) + 1e-3)
        et = np.abs(dt[:, diag] - want_t[:, diag]).max(1) / (np.abs(want_t[:, diag]).max(1) + 1e-2)
        ez = np.abs(dz - want_z) / (np.abs(want_z) + 1e-1)
        ew = ew / np.where(shear, 5.0, 1.0)   # the atomics path carries the clipped pixels' cancellation noise
        bad = []
        if ew.max() >= 5e-5: bad.append((name, "dU", int(ew.argmax()), float(ew.max()), thi[ew.argmax()]))
        if et.max() >= 2e-4: bad.append((name, "dtheta", int(et.argmax()), float(et.max()), thi[et.argmax()], dt[et.argmax()], want_t[et.argmax()]))
        if ez.max() >= 5e-5: bad.append((name, "dz", int(ez.argmax()), float(ez.max()), thi[ez.argmax()]))
        assert not bad, bad
        # fp32 oracle, parity tolerance, model regime
        assert relnorm(dw[model_rows], f32_w[model_rows]) < 1e-4 and relnorm(dz[model_rows], f32_z[model_rows]) < 1e-4, name
        assert relnorm(dt[model_rows][:, diag], f32_t[model_rows][:, diag]) < 1e-4, name
    # off-diagonal entries: exact zeros from the axis kernel, the real gradient on the shear row and in the staged kernel
    dt_axis, dt_staged = res[1][1], res[0][1]
    assert not dt_axis[~shear][:, [1, 3]].any()
    assert relnorm(dt_axis[shear], want_t[shear]) < 1e-4
    assert relnorm(dt_staged[:, [1, 3]], want_t[:, [1, 3]]) < 1e-4
    dead = stop >= 0.99
    assert not res[1][0][dead].any() and not res[1][1][dead].any() and not res[1][2][dead].any()


@pytest.mark.parametrize("T,B,with_canvas", [(1, 7, True), (3, 33, False), (3, 64, True), (5, 10, False), (30, 3, True)])
def test_writeback_steps_bit_identical_to_sequential(T, B, with_canvas):
    """air_st_writeback_canvas_fwd_steps == T consecutive single-step write-backs, bit for bit (T = 30 exceeds the
    shared-memory budget of the fused kernel and takes the documented fallback)."""
    rng = np.random.RandomState(70 + T + B)
    NF = 20
    win = cu(rng.rand(T, B, 784).astype(np.float32))
    thi = np.zeros((T, B, 6), np.float32)
    for t in range(T):
        thi[t] = air_thetas(rng, B)[1].reshape(B, 6)
    thi[0, 0] = [1.2, 0.3, 0.1, -0.2, 1.1, 0.0]                       # one sheared row: per-pixel path
    thi = cu(thi)
    fields = torch.zeros(T, NF, B, device=DEV)
    fields[:, 9] = cu(rng.rand(T, B).astype(np.float32))              # z
    fields[:, 15] = cu(rng.choice(np.array([0.0, 0.5, 1.3], np.float32), (T, B)))   # stop_new
    canvas0 = cu(rng.rand(B, 2500).astype(np.float32)) if with_canvas else None
    seq = canvas0.clone() if with_canvas else torch.empty(B, 2500, device=DEV)
    for t in range(T):
        ops.writeback_canvas_fwd(win[t], thi[t], fields[t, 9], fields[t, 15], 0.99,
                                 seq if (t > 0 or with_canvas) else None, seq, 28, 28, 50, 50)
    fused = torch.full((B, 2500), 7.0, device=DEV)
    ops.writeback_canvas_fwd_steps(win, thi, fields[0, 9], fields[0, 15], NF * B, 0.99, canvas0, fused, 28, 28, 50, 50)
    assert torch.equal(fused, seq)
    if with_canvas:                                                   # in place
        inpl = canvas0.clone()
        ops.writeback_canvas_fwd_steps(win, thi, fields[0, 9], fields[0, 15], NF * B, 0.99, inpl, inpl, 28, 28, 50, 50)
        assert torch.equal(inpl, seq)


@pytest.mark.parametrize("T,B", [(3, 64), (5, 12), (2, 7), (1, 8)])
def test_steps_launches_bit_identical_to_per_step_calls(T, B):
    """air_st_forward_steps / air_st_backward_steps / air_st_writeback_canvas_bwd_steps == T single-step calls, bit for
    bit (B = 7 is not a multiple of 4: the crop forward takes the documented per-step fallback)."""
    rng = np.random.RandomState(90 + T + B)
    NF = 20
    U = cu(rng.rand(B, 2500).astype(np.float32))
    th = np.zeros((T, B, 6), np.float32); thi = np.zeros((T, B, 6), np.float32)
    for t in range(T):
        a, b_ = air_thetas(rng, B)
        th[t], thi[t] = a.reshape(B, 6), b_.reshape(B, 6)
    thi[0, 1] = [1.3, 0.2, 0.1, -0.1, 1.2, 0.0]                               # a sheared row (atomics fallback inside)
    th, thi = cu(th), cu(thi)
    # crop forward
    win_b, win_s = torch.empty(T, B, 784, device=DEV), torch.empty(T, B, 784, device=DEV)
    ops.st_forward_steps(U, th, win_b, 50, 50, 1, 28, 28)
    for t in range(T):
        ops.st_forward(U, th[t], win_s[t], 50, 50, 1, 28, 28)
    assert torch.equal(win_b, win_s)
    # crop backward
    dwin = cu(rng.randn(T, B, 784).astype(np.float32))
    dth_b, dth_s = torch.empty(T, B, 6, device=DEV), torch.empty(T, B, 6, device=DEV)
    ops.st_backward_steps(U, th, dwin, dth_b, 50, 50, 1, 28, 28)
    for t in range(T):
        ops.st_backward(U, th[t], dwin[t], None, dth_s[t], 50, 50, 1, 28, 28)
    assert torch.equal(dth_b, dth_s)
    # fused write-back backward against one dcanvas
    fields = torch.zeros(T, NF, B, device=DEV)
    fields[:, 9] = cu(rng.rand(T, B).astype(np.float32) * 0.9 + 0.05)
    fields[:, 15] = cu(rng.choice(np.array([0.0, 0.5, 1.3], np.float32), (T, B)))
    dcanvas = cu(rng.randn(B, 2500).astype(np.float32))
    recon = cu(rng.rand(T, B, 784).astype(np.float32))
    for sig, axis in ((True, True), (False, True), (False, False)):
        outs = []
        for batched in (True, False):
            dw, dt, dz = torch.full((T, B, 784), 7.0, device=DEV), torch.full((T, B, 6), 7.0, device=DEV), torch.full((T, B), 7.0, device=DEV)
            if batched:
                ops.writeback_canvas_bwd_steps(recon, thi, fields[0, 9], fields[0, 15], NF * B, 0.99, dcanvas, dw, dt, dz,
                                               28, 28, 50, 50, window_is_sigmoid=sig, axis_aligned_theta=axis)
            else:
                for t in range(T):
                    ops.writeback_canvas_bwd(recon[t], thi[t], fields[t, 9], fields[t, 15], 0.99, dcanvas, dw[t], dt[t], dz[t],
                                             28, 28, 50, 50, window_is_sigmoid=sig, axis_aligned_theta=axis)
            outs.append((dw, dt, dz))
        for a, b_ in zip(*outs):
            # the sheared row's dU goes through shared-memory float atomics (order not fixed): close, not identical
            assert torch.allclose(a[0, 1], b_[0, 1], rtol=1e-4, atol=1e-5), (sig, axis)
            a, b_ = a.clone(), b_.clone()
            a[0, 1] = 0; b_[0, 1] = 0
            assert torch.equal(a, b_), (sig, axis)


def test_concrete_step_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "concrete.npz"))
    for train in (0, 1):
        y, z, zp, kl, sn, ln, dn = ab.concrete_step(cu(g["log_odds"]), cu(g["u"]), cu(g["stop_prev"]),
                                                    cu(g["loss_prev"]), cu(g["digits_prev"]), float(g["prior"]),
                                                    float(g["temperature"]), float(g["thr"]), bool(train))
        t = f"_train{train}"
        assert np.array_equal(dn.cpu().numpy(), g["digits_new" + t])                      # counts: exact
        assert np.array_equal((sn.cpu().numpy() < 0.99), (g["stop_new" + t] < 0.99))      # masks: exact
        if not train:
            assert np.array_equal(z.cpu().numpy(), g["z" + t])                            # rounded z: exact
        for name, v in (("y", y), ("z", z), ("z_prob", zp), ("kl", kl), ("stop_new", sn), ("loss_new", ln)):
            np.testing.assert_allclose(v.cpu().numpy(), g[name + t], rtol=1e-5, atol=2e-6, err_msg=name)


def test_concrete_step_backward():
    rng = np.random.RandomState(51)
    n = 512
    lo = (rng.randn(n) * 2).astype(np.float32)
    u = rng.rand(n).astype(np.float32)
    stop_prev = rng.choice(np.array([0.0, 0.5, 1.2], np.float32), n)
    lt = torch.from_numpy(lo).requires_grad_(True)
    y = O.concrete_binary_pre_sigmoid_sample(lt, 0.8, torch.from_numpy(u))
    z = torch.sigmoid(y)
    kl = O.concrete_binary_kl_mc_sample(y, -1.5, 0.8, lt, 0.8)
    gz, gk = rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32)
    loss = torch.where(torch.from_numpy(stop_prev) < 0.99, kl, torch.zeros_like(kl))
    (z * torch.from_numpy(gz)).sum().add((loss * torch.from_numpy(gk)).sum()).backward()
    lg = cu(lo).requires_grad_(True)
    zf = torch.zeros(n, device=DEV)
    _, zz, _, _, _, ln, _ = ab.concrete_step(lg, cu(u), cu(stop_prev), zf, torch.zeros(n, device=DEV, dtype=torch.int32),
                                             -1.5, 0.8, 0.99, True)
    ((zz * cu(gz)).sum() + (ln * cu(gk)).sum()).backward()
    assert relnorm(lg.grad.cpu().numpy(), lt.grad.numpy()) < 1e-5


def test_reference_signature_concrete_functions():
    rng = np.random.RandomState(52)
    lo, u = (rng.randn(100) * 2).astype(np.float32), rng.rand(100).astype(np.float32)
    y = ab.concrete_binary_pre_sigmoid_sample(cu(lo), 0.7, u=cu(u))
    want = O.concrete_binary_pre_sigmoid_sample(torch.from_numpy(lo), 0.7, torch.from_numpy(u))
    np.testing.assert_allclose(y.cpu().numpy(), want.numpy(), rtol=1e-5, atol=2e-6)
    kl = ab.concrete_binary_kl_mc_sample(y, -2.0, 0.7, cu(lo), 0.7)
    np.testing.assert_allclose(kl.cpu().numpy(), O.concrete_binary_kl_mc_sample(want, -2.0, 0.7, torch.from_numpy(lo), 0.7).numpy(),
                               rtol=1e-4, atol=1e-5)
    yy, sig = ab.concrete_binary_sample(cu(lo), 0.7, hard=True, u=cu(u))
    wy, wsig = O.concrete_binary_sample(torch.from_numpy(lo), 0.7, torch.from_numpy(u), hard=True)
    np.testing.assert_allclose(yy.cpu().numpy(), wy.numpy(), rtol=1e-5, atol=1e-5)
    assert np.array_equal(sig.cpu().numpy(), wsig.numpy())


def test_full_size_properties_b65536():
    """BASELINE config sizes: size-independent properties instead of the (slow) oracle."""
    B = 65536
    gen = torch.Generator(device=DEV).manual_seed(0)
    U = torch.rand(B, 50, 50, 1, device=DEV, generator=gen)
    s = torch.rand(B, device=DEV, generator=gen) * 0.6 + 0.3
    xy = torch.rand(B, 2, device=DEV, generator=gen) - 0.5
    th = torch.zeros(B, 2, 3, device=DEV)
    th[:, 0, 0] = s; th[:, 1, 1] = s; th[:, :, 2] = xy
    out = ab.transformer(U, th, (28, 28))
    # (1) same rows computed in a small batch are bit-identical (no cross-image state)
    idx = torch.tensor([0, 1, 4097, 65535], device=DEV)
    assert torch.equal(ab.transformer(U[idx].contiguous(), th[idx].contiguous(), (28, 28)), out[idx])
    # (2) linearity in U: ST(2U) == 2 ST(U) exactly (power-of-two scaling commutes with rounding)
    assert torch.equal(ab.transformer(U * 2, th, (28, 28)), out * 2)
    # (3) convex combination: outputs inside [min U, max U] up to rounding for in-range windows
    assert out.min() >= -1e-5 and out.max() <= 1 + 1e-5
    # (4) oracle spot check on a strided sample
    sub = torch.arange(0, B, 1024, device=DEV)
    want = C.st_forward(U[sub].cpu().numpy(), th[sub].cpu().numpy(), (28, 28))
    assert np.array_equal(out[sub].cpu().numpy(), want)


def test_full_size_properties_fused_and_backward_b65536():
    """B = 65536 (BASELINE configs[1]): properties that need no oracle."""
    B = 65536
    gen = torch.Generator(device=DEV).manual_seed(1)
    r = lambda *s: torch.rand(*s, device=DEV, generator=gen)
    s = r(B) * 0.6 + 0.3
    xy = r(B, 2) - 0.5
    th = torch.zeros(B, 6, device=DEV)
    th[:, 0] = s; th[:, 4] = s; th[:, 2] = xy[:, 0]; th[:, 5] = xy[:, 1]
    thi = torch.zeros(B, 6, device=DEV)
    thi[:, 0] = 1 / s; thi[:, 4] = 1 / s; thi[:, 2] = -xy[:, 0] / s; thi[:, 5] = -xy[:, 1] / s
    win, z, canvas = r(B, 28, 28), r(B), r(B, 50, 50)
    stop = (r(B) > 0.7).float() * 1.5
    L, c = ab._cabi.lib(), ab._cabi
    out = torch.empty_like(canvas)
    c.check(L.air_st_writeback_canvas_fwd(c.ptr(win), c.ptr(thi), c.ptr(z), c.ptr(stop), 0.99, c.ptr(canvas), c.ptr(out),
                                          B, 28, 28, 50, 50, c.stream()), "fwd")
    # (1) in place == out of place, bit for bit
    inpl = canvas.clone()
    c.check(L.air_st_writeback_canvas_fwd(c.ptr(win), c.ptr(thi), c.ptr(z), c.ptr(stop), 0.99, c.ptr(inpl), c.ptr(inpl),
                                          B, 28, 28, 50, 50, c.stream()), "fwd inplace")
    assert torch.equal(inpl, out)
    # (1b) canvas_in == NULL stands for an all-zero canvas (first loop step), bit for bit
    zero_out, null_out = torch.empty_like(canvas), torch.empty_like(canvas)
    zc = torch.zeros_like(canvas)
    c.check(L.air_st_writeback_canvas_fwd(c.ptr(win), c.ptr(thi), c.ptr(z), c.ptr(stop), 0.99, c.ptr(zc), c.ptr(zero_out),
                                          B, 28, 28, 50, 50, c.stream()), "fwd zero canvas")
    c.check(L.air_st_writeback_canvas_fwd(c.ptr(win), c.ptr(thi), c.ptr(z), c.ptr(stop), 0.99, None, c.ptr(null_out),
                                          B, 28, 28, 50, 50, c.stream()), "fwd null canvas")
    assert torch.equal(zero_out, null_out)
    # (2) stopped rows are untouched; live rows equal canvas + z * ST(window) computed by the plain ST kernel
    dead = stop >= 0.99
    assert torch.equal(out[dead], canvas[dead])
    plain = ab.transformer(win.unsqueeze(3), thi, (50, 50))[..., 0]
    want = canvas + torch.where(~dead[:, None, None], z[:, None, None] * plain, torch.zeros_like(plain))
    assert torch.equal(out, want)
    # (3) backward: batch-composition invariance and exact linearity in the upstream gradient
    g = torch.randn(B, 50, 50, device=DEV, generator=gen)
    def bwd(wi, ti, zi, si, gi):
        n = wi.shape[0]
        dw, dt, dz = torch.empty_like(wi), torch.empty(n, 6, device=DEV), torch.empty(n, device=DEV)
        c.check(L.air_st_writeback_canvas_bwd(c.ptr(wi), c.ptr(ti), c.ptr(zi), c.ptr(si), 0.99, c.ptr(gi), c.ptr(dw),
                                              c.ptr(dt), c.ptr(dz), 0, n, 28, 28, 50, 50, c.stream()), "bwd")
        return dw, dt, dz
    dw, dt, dz = bwd(win, thi, z, stop, g)
    idx = torch.tensor([0, 5, 4097, 40000, 65535], device=DEV)
    sub = bwd(win[idx].contiguous(), thi[idx].contiguous(), z[idx].contiguous(), stop[idx].contiguous(), g[idx].contiguous())
    assert torch.equal(sub[0], dw[idx]) and torch.equal(sub[1], dt[idx]) and torch.equal(sub[2], dz[idx])
    dw2, dt2, dz2 = bwd(win, thi, z, stop, g * 4)
    assert torch.equal(dw2, dw * 4) and torch.equal(dt2, dt * 4) and torch.equal(dz2, dz * 4)
    assert not dw[dead].any() and not dt[dead].any() and not dz[dead].any()
    # (3b) the warp-specialised axis-aligned kernel: same invariances, and agreement with the staged kernel
    def bwd_axis(wi, ti, zi, si, gi):
        n = wi.shape[0]
        dw_, dt_, dz_ = torch.empty_like(wi), torch.empty(n, 6, device=DEV), torch.empty(n, device=DEV)
        c.check(L.air_st_writeback_canvas_bwd(c.ptr(wi), c.ptr(ti), c.ptr(zi), c.ptr(si), 0.99, c.ptr(gi), c.ptr(dw_),
                                              c.ptr(dt_), c.ptr(dz_), 2, n, 28, 28, 50, 50, c.stream()), "bwd axis")
        return dw_, dt_, dz_
    adw, adt, adz = bwd_axis(win, thi, z, stop, g)
    asub = bwd_axis(win[idx].contiguous(), thi[idx].contiguous(), z[idx].contiguous(), stop[idx].contiguous(), g[idx].contiguous())
    assert torch.equal(asub[0], adw[idx]) and torch.equal(asub[1], adt[idx]) and torch.equal(asub[2], adz[idx])
    adw2, adt2, adz2 = bwd_axis(win, thi, z, stop, g * 4)
    assert torch.equal(adw2, adw * 4) and torch.equal(adt2, adt * 4) and torch.equal(adz2, adz * 4)
    assert not adw[dead].any() and not adt[dead].any() and not adz[dead].any()
    rel = lambda a, b_: float((a - b_).norm() / b_.norm())
    assert rel(adw, dw) < 2e-6 and rel(adz, dz) < 2e-6 and rel(adt[:, [0, 2, 4, 5]], dt[:, [0, 2, 4, 5]]) < 2e-5
    assert not adt[:, [1, 3]].any()
    # (4) dz is the inner product <dU / z, U> of its own outputs (the identity the kernel uses), and also
    #     equals sum(g * ST(window)) computed independently, to fp32 accuracy
    live = ~dead
    ref_dz = (g * plain).flatten(1).sum(1)
    assert torch.allclose(dz[live], ref_dz[live], rtol=2e-4, atol=2e-3)
    # (5) crop backward: subset invariance + linearity
    U = r(B, 50, 50, 1)
    gw = torch.randn(B, 28, 28, 1, device=DEV, generator=gen)
    def crop_bwd(Ui, ti, gi):
        n = Ui.shape[0]
        dt_ = torch.empty(n, 6, device=DEV)
        c.check(L.air_st_backward(c.ptr(Ui), c.ptr(ti), c.ptr(gi), None, c.ptr(dt_), n, 50, 50, 1, 28, 28, c.stream()), "crop bwd")
        return dt_
    d1 = crop_bwd(U, th, gw)
    assert torch.equal(crop_bwd(U[idx].contiguous(), th[idx].contiguous(), gw[idx].contiguous()), d1[idx])
    assert torch.equal(crop_bwd(U, th, gw * 0.5), d1 * 0.5)

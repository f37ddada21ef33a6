"""GPU unit parity tests of the dense / elementwise kernels behind the C ABI."""
import numpy as np
import pytest
import torch

import air_b200 as ab
from air_b200 import ops
from oracle import air_oracle as O
from oracle import c_oracle as C
from tests.parity_util import relnorm

pytestmark = pytest.mark.gpu
DEV = "cuda"
K = ab._cabi


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("M,N,Kd", [(64, 64, 64), (130, 70, 33), (257, 320, 256), (5, 7, 3), (300, 100, 50),
                                    (1, 1, 1), (128, 1024, 2500), (2500, 1024, 64)])
@pytest.mark.parametrize("tA,tB", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_exact_is_k_sequential_fma(M, N, Kd, tA, tB):
    rng = np.random.RandomState(M + N + Kd)
    A = rng.randn(M, Kd).astype(np.float32)
    Bm = rng.randn(Kd, N).astype(np.float32)
    Ci = rng.randn(M, N).astype(np.float32)
    bias = rng.randn(N).astype(np.float32)
    want = C.gemm_seq_fma(A, Bm, Ci, bias)
    At = cu(A.T.copy()) if tA else cu(A)
    Bt = cu(Bm.T.copy()) if tB else cu(Bm)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(At, Bt, out, Cinit=cu(Ci), bias=cu(bias), tA=tA, tB=tB)
    assert np.array_equal(out.cpu().numpy(), want)
    # accumulate in place (Cinit aliases out) and no bias
    out2 = cu(Ci).clone()
    ops.gemm(At, Bt, out2, Cinit=out2, tA=tA, tB=tB)
    assert np.array_equal(out2.cpu().numpy(), C.gemm_seq_fma(A, Bm, Ci, None))


def test_gemm_strided_views_and_epilogues():
    rng = np.random.RandomState(0)
    big = cu(rng.randn(40, 100).astype(np.float32))
    A = big[:, 10:60]                       # [40,50] view, lda = 100
    W = cu(rng.randn(50, 24).astype(np.float32))
    b = cu(rng.randn(24).astype(np.float32))
    pre = torch.from_numpy(C.gemm_seq_fma(A.cpu().numpy(), W.cpu().numpy(), None, b.cpu().numpy()))
    for epi, f in ((K.EPI_RELU, torch.relu), (K.EPI_SOFTPLUS, O.tf_softplus), (K.EPI_NONE, lambda v: v)):
        out = torch.empty(40, 24, device=DEV)
        ops.gemm(A, W, out, bias=b, epi=epi)
        np.testing.assert_allclose(out.cpu().numpy(), f(pre).numpy(), rtol=2e-6, atol=1e-7)
    # backward epilogues: dX = (dY W^T) * act'(aux)
    dY = cu(rng.randn(40, 24).astype(np.float32))
    x_pre = torch.from_numpy(rng.randn(40, 50).astype(np.float32) * 6)
    for epi, post, dact in ((K.EPI_MUL_DRELU, torch.relu(x_pre), (x_pre > 0).float()),
                            (K.EPI_MUL_DSOFTPLUS, O.tf_softplus(x_pre), torch.sigmoid(x_pre))):
        dX = torch.empty(40, 50, device=DEV)
        ops.gemm(dY, W, dX, tB=True, aux=post.to(DEV), epi=epi)
        want = (dY.cpu() @ W.cpu().t()) * dact
        assert relnorm(dX, want) < 2e-6


def test_softplus_backward_tiny_activations():
    """softplus' is taken from the stored OUTPUT y: sigmoid(x) = 1 - exp(-y) = -expm1(-y), which
    stays accurate for very negative inputs.  TF's forward log(exp(x) + 1) itself carries up to
    ~1e-3 relative rounding for x ~ -10 (1 + 4.5e-5 in fp32), and the derivative inherits exactly
    that; elsewhere it is accurate to a few ulp."""
    x = torch.tensor([[-20.0, -14.0, -10.0, -3.0, 0.0, 5.0, 14.5, 30.0]])
    post = O.tf_softplus(x)
    dX = torch.empty(1, 8, device=DEV)
    ops.gemm(torch.ones(1, 1, device=DEV), torch.ones(8, 1, device=DEV), dX, tB=True, aux=post.to(DEV),
             epi=K.EPI_MUL_DSOFTPLUS)
    got, want = dX.cpu().numpy()[0], torch.sigmoid(x.double()).numpy()[0]
    np.testing.assert_allclose(got[[0, 1, 4, 5, 6, 7]], want[[0, 1, 4, 5, 6, 7]], rtol=3e-6)
    np.testing.assert_allclose(got[[2, 3]], want[[2, 3]], rtol=2e-3)
    np.testing.assert_allclose(got, -np.expm1(-post.double().numpy()[0]), rtol=3e-6)   # exact w.r.t. the stored y


def test_gemm_errors():
    a = torch.zeros(4, 4, device=DEV)
    with pytest.raises(ab.AirError):
        ops.gemm(a, torch.zeros(5, 4, device=DEV), torch.zeros(4, 4, device=DEV))
    with pytest.raises(ab.AirError, match="needs aux"):
        ops.gemm(a, a, torch.zeros(4, 4, device=DEV), epi=K.EPI_MUL_DRELU)


@pytest.mark.parametrize("mode", [0, 1])
def test_gemm_k0_is_cinit_plus_bias(mode):
    """K == 0 (the first LSTM step: zero initial state, air_model.py:540-542): out = epi(Cinit + bias), no operands read."""
    torch.manual_seed(3)
    h0, Kh = torch.zeros(128, 256, device=DEV), torch.randn(256, 1024, device=DEV)
    cinit, bias = torch.randn(128, 1024, device=DEV), torch.randn(1024, device=DEV)
    out = torch.full((128, 1024), 7.0, device=DEV)
    ops.gemm(h0[:, :0], Kh[:0], out, Cinit=cinit, bias=bias, mode=mode)
    assert torch.equal(out, cinit + bias)
    full = torch.empty_like(out)
    ops.gemm(h0, Kh, full, Cinit=cinit, bias=bias, mode=0)     # fma(0, k, acc) == acc: same bits as the full GEMM
    assert torch.equal(out, full)


def test_lstm_pointwise_fwd_bwd():
    torch.manual_seed(0)
    B, H = 37, 256
    gates = torch.randn(B, 4 * H, requires_grad=True)
    c0 = torch.randn(B, H, requires_grad=True)
    gi, gj, gf, go = torch.split(gates, H, dim=1)
    c = c0 * torch.sigmoid(gf + 1.0) + torch.sigmoid(gi) * torch.tanh(gj)
    h = torch.tanh(c) * torch.sigmoid(go)
    dh, dc = torch.randn(B, H), torch.randn(B, H)
    (h * dh).sum().add((c * dc).sum()).backward()
    cg, hg = torch.empty(B, H, device=DEV), torch.empty(B, H, device=DEV)
    ops.lstm_fwd(gates.detach().to(DEV), c0.detach().to(DEV), cg, hg)
    np.testing.assert_allclose(cg.cpu().numpy(), c.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(hg.cpu().numpy(), h.detach().numpy(), rtol=1e-5, atol=1e-6)
    dg, dcp = torch.empty(B, 4 * H, device=DEV), torch.empty(B, H, device=DEV)
    dsum = torch.ones(B, 4 * H, device=DEV)
    ops.lstm_bwd(gates.detach().to(DEV), c0.detach().to(DEV), cg, dh.to(DEV), dc.to(DEV), dg, dcp, dsum)
    assert relnorm(dg, gates.grad) < 1e-5 and relnorm(dcp, c0.grad) < 1e-5
    assert relnorm(dsum - 1.0, gates.grad) < 1e-5
    # zero initial state (c_prev = NULL)
    ops.lstm_fwd(gates.detach().to(DEV), None, cg, hg)
    c_z = torch.sigmoid(gi) * torch.tanh(gj)
    np.testing.assert_allclose(cg.cpu().numpy(), c_z.detach().numpy(), rtol=1e-5, atol=1e-6)


def test_bce_loss_and_grad():
    rng = np.random.RandomState(1)
    B, N = 9, 2500
    canvas = (rng.rand(B, N).astype(np.float32) * 1.6 - 0.3)
    canvas[0, :5] = [0.0, 1.0, -0.0, 1.0000001, -1e-8]           # tie rules at the clip boundaries
    x = (rng.rand(B, N) > 0.8).astype(np.float32) * rng.rand(B, N).astype(np.float32)
    ct = torch.from_numpy(canvas).requires_grad_(True)
    r = torch.clamp(ct, 0.0, 1.0)
    xt = torch.from_numpy(x)
    rl = -torch.sum(xt * torch.log(r + O.EPS) + (1.0 - xt) * torch.log(1.0 - r + O.EPS), 1)
    (rl.sum() * 0.125).backward()
    recon, loss, dcv = torch.empty(B, N, device=DEV), torch.empty(B, device=DEV), torch.empty(B, N, device=DEV)
    ops.bce_loss(cu(canvas), cu(x), recon, loss, dcv, 0.125)
    assert np.array_equal(recon.cpu().numpy(), r.detach().numpy())
    np.testing.assert_allclose(loss.cpu().numpy(), rl.detach().numpy(), rtol=2e-6)
    assert relnorm(dcv, ct.grad) < 1e-6
    assert np.array_equal(dcv.cpu().numpy() != 0, ct.grad.numpy() != 0)   # identical pass/block pattern


def test_colsum_deterministic():
    rng = np.random.RandomState(2)
    for B, N in ((4096, 1024), (100, 7), (33, 320), (1, 50)):
        X = rng.randn(B, N).astype(np.float32)
        ws = torch.zeros(int(K.lib().air_colsum_workspace(B, N)), device=DEV)
        out = torch.full((N,), 3.0, device=DEV)
        ops.colsum(cu(X), out, False, ws)
        np.testing.assert_allclose(out.cpu().numpy(), X.astype(np.float64).sum(0), rtol=1e-5, atol=1e-4)
        first = out.clone()
        ops.colsum(cu(X), out, True, ws)
        np.testing.assert_allclose(out.cpu().numpy(), 2 * first.cpu().numpy(), rtol=1e-6)
        again = torch.empty(N, device=DEV)
        ops.colsum(cu(X), again, False, ws)
        assert torch.equal(again, first)


def test_colsum_multi_and_reduce_rows():
    torch.manual_seed(5)
    shapes = [(12288, 512), (12288, 100), (4096, 1024), (37, 7), (12288, 784)]
    Xs = [torch.randn(r, n + 3, device=DEV)[:, :n] for r, n in shapes]          # row-strided views
    outs = [torch.full((n,), 3.0, device=DEV) for _, n in shapes]
    accs = [False, True, False, False, True]
    items = ops.colsum_items(list(zip(Xs, outs, accs)))
    ws = torch.zeros(ops.colsum_multi_workspace(items), device=DEV)
    ops.colsum_multi(items, ws)
    for X, o, a in zip(Xs, outs, accs):
        want = X.double().sum(0) + (3.0 if a else 0.0)
        assert relnorm(o, want) < 2e-6
    first = [o.clone() for o in outs]
    for o in outs:
        o.fill_(3.0)
    ops.colsum_multi(items, ws)
    assert all(torch.equal(a, b) for a, b in zip(first, outs))                       # deterministic
    # the single-tensor entry point gives the same bits
    single = torch.empty(512, device=DEV)
    ops.colsum(Xs[0], single, False, torch.zeros(int(K.lib().air_colsum_workspace(12288, 512)), device=DEV))
    assert torch.equal(single, outs[0])
    part = torch.randn(300, 455, device=DEV)
    out = torch.empty(448, device=DEV)
    ops.reduce_rows(part, 300, 455, 448, out)
    assert relnorm(out, part[:, :448].double().sum(0)) < 2e-6
    out7 = torch.ones(7, device=DEV)
    ops.reduce_rows(part.view(-1)[448:], 300, 455, 7, out7, accumulate=True)
    assert relnorm(out7, part[:, 448:].double().sum(0) + 1.0) < 2e-6


def test_device_canvas_generator_bit_exact_and_shardable():
    from air_b200 import data
    from oracle import synth_oracle as S
    im, cnt, pos, box = data.device_canvases(193, seed=11, first_index=5, with_boxes=True)
    want_im, want_cnt, want_pos, want_box = S.synth_canvases(193, seed=11, first_index=5, with_boxes=True)
    assert np.array_equal(cnt.cpu().numpy(), want_cnt) and np.array_equal(im.cpu().numpy(), want_im)   # bit-exact
    assert np.array_equal(pos.cpu().numpy().reshape(193, -1), want_pos) and np.array_equal(box.cpu().numpy().reshape(193, -1), want_box)
    # the generator's invariants (multi_mnist.py:141-166): boxes inside the canvas, label == boxes written, the ink of a
    # canvas is exactly the union of its boxes' ink (no pixel overlap), nothing outside the boxes
    P, Bx, I = pos.cpu().numpy(), box.cpu().numpy(), im.cpu().numpy().reshape(193, 50, 50)
    for b in range(193):
        k = int(cnt[b])
        assert (Bx[b, :k] > 0).all() and (Bx[b, k:] == 0).all()
        mask = np.zeros((50, 50), bool)
        for (x, y), (w, h) in zip(P[b, :k], Bx[b, :k]):
            assert 0 <= x and x + w <= 50 and 0 <= y and y + h <= 50
            mask[y:y + h, x:x + w] = True
        assert not I[b][~mask].any()
    # image i of a seed does not depend on the batch it is generated in (data-parallel shards)
    a, ca = data.device_canvases(64, seed=11, first_index=5)
    b, cb = data.device_canvases(129, seed=11, first_index=69)
    assert torch.equal(torch.cat([a, b]), im) and torch.equal(torch.cat([ca, cb]), cnt)
    # statistics of the stand-in data set (multi_mnist.py strata: 0 / 1 / 2 digits equally likely)
    big, cbig = data.device_canvases(30000, seed=1)
    frac = torch.bincount(cbig, minlength=3).float() / 30000
    assert (frac - 1 / 3).abs().max() < 0.02 and big.min() == 0 and 0.9 < big.max() <= 1.0
    empty = cbig == 0
    assert not big[empty].any() and (big[~empty] > 0).any(1).all()
    # in place refill of a model's input buffers
    buf = (torch.empty(8, 2500, device=DEV), torch.empty(8, device=DEV, dtype=torch.int32))
    data.device_canvases(8, seed=11, first_index=5, out=buf)
    assert torch.equal(buf[0], im[:8])


def test_adam_step_matches_tf_semantics():
    torch.manual_seed(3)
    orc = O.AIROracle(seed=1)
    st = ab.ParamStore(DEV, 2500, 784, 256, 64, 50, (512, 256), (256, 512), seed=1)
    st.state[4] = 1e-4
    ws = torch.zeros(int(K.lib().air_adam_workspace(st.n)), device=DEV)
    for step in range(3):
        grads = {k: torch.randn_like(v) * (10.0 if step == 0 else 1e-3) for k, v in orc.params.items()}
        for k, gview in st.named_grads().items():
            gview.copy_(grads[k].to(DEV))
        clipped, norm = O.AIROracle.clip_by_global_norm(grads, 1.0)
        orc.adam_apply(clipped)
        ops.adam_step(st.flat, st.grad, st.adam_m, st.adam_v, st.state, 1.0, 0.9, 0.999, 1e-8, 1.0, ws)
        assert abs(st.state[3].item() - norm.item()) / norm.item() < 1e-5
        for k, v in st.named_views().items():
            np.testing.assert_allclose(v.cpu().numpy(), orc.params[k].numpy(), rtol=2e-5, atol=2e-8, err_msg=k)
    assert st.global_step == 3
    np.testing.assert_allclose(st.state[:2].cpu().numpy(), [0.9 ** 4, 0.999 ** 4], rtol=1e-6)


def test_adam_step_skips_a_non_finite_gradient_only_when_asked():
    """AIR_ADAM_SKIP_NONFINITE: an overflowed gradient (inf - inf in the un-cancelled corner products of a collapsed window)
    leaves parameters, Adam slots, beta powers and global_step untouched and is counted; without the flag the step does
    what the reference's clip_by_global_norm + ApplyAdam do -- NaN everywhere."""
    st = ab.ParamStore(DEV, 2500, 784, 256, 64, 50, (512, 256), (256, 512), seed=2)
    st.state[4] = 1e-4
    ws = torch.zeros(int(K.lib().air_adam_workspace(st.n)), device=DEV)
    st.grad.normal_()
    ops.adam_step(st.flat, st.grad, st.adam_m, st.adam_v, st.state, 1.0, 0.9, 0.999, 1e-8, 1.0, ws, skip_nonfinite=True)
    before = (st.flat.clone(), st.adam_m.clone(), st.adam_v.clone(), st.state.clone())
    assert st.global_step == 1 and st.state[5].item() == 0
    for bad in (float("inf"), float("nan")):
        st.grad.normal_()
        st.grad[12345] = bad
        ops.adam_step(st.flat, st.grad, st.adam_m, st.adam_v, st.state, 1.0, 0.9, 0.999, 1e-8, 1.0, ws, skip_nonfinite=True)
        assert torch.equal(st.flat, before[0]) and torch.equal(st.adam_m, before[1]) and torch.equal(st.adam_v, before[2])
        assert torch.equal(st.state[:3], before[3][:3]) and not torch.isfinite(st.state[3])
    assert st.state[5].item() == 2 and st.global_step == 1
    st.grad.normal_()   # a finite gradient afterwards: business as usual
    ops.adam_step(st.flat, st.grad, st.adam_m, st.adam_v, st.state, 1.0, 0.9, 0.999, 1e-8, 1.0, ws, skip_nonfinite=True)
    assert st.global_step == 2 and torch.isfinite(st.flat).all() and not torch.equal(st.flat, before[0])
    st.grad[7] = float("nan")
    ops.adam_step(st.flat, st.grad, st.adam_m, st.adam_v, st.state, 1.0, 0.9, 0.999, 1e-8, 1.0, ws)
    assert not torch.isfinite(st.flat).any() and st.global_step == 3   # the reference's behaviour: the NaN norm reaches every variable


def test_anneal_kernel():
    st = torch.zeros(8, device=DEV)
    out = torch.zeros(1, device=DEV)
    sched = O.DEFAULT_ANNEALING["z_pres_prior_log_odds"]
    for step in (0, 1, 2999, 3000, 10000, 60000):
        st[2] = float(step)
        ops.anneal(st, sched, out)
        assert out.item() == pytest.approx(O.annealed_value(sched, step).item(), rel=2e-5, abs=1e-6)


def test_vae_function_reference_signature():
    torch.manual_seed(4)
    B = 33
    x = torch.rand(B, 784)
    n1, n2 = torch.randn(B, 50), torch.randn(B, 784)
    xg = x.to(DEV).requires_grad_(True)
    rec, mean, logvar, latent = ab.vae(xg, 784, (512, 256), 50, (256, 512), 0.3, scope="t_vae", seed=5,
                                       noise_latent=n1.to(DEV), noise_like=n2.to(DEV))
    assert latent is mean                                    # vae.py:43 returns the mean twice
    store = ab.air.vae.variables("t_vae")[0] if hasattr(ab, "air") else None
    from importlib import import_module
    store = import_module("tf-attend-infer-repeat_b200.air.vae").variables("t_vae")[0]
    params = {"vae/" + k[4:]: v.cpu() for k, v in store.named_views().items() if k.startswith("vae/")}
    xt = x.clone().requires_grad_(True)
    orec, omean, olv, _ = O.vae(xt, params, "vae/", 2, 2, n1, n2, 0.3)
    np.testing.assert_allclose(rec.detach().cpu().numpy(), orec.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(mean.detach().cpu().numpy(), omean.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(logvar.detach().cpu().numpy(), olv.detach().numpy(), rtol=1e-5, atol=1e-6)
    # a caller-side loss that uses all outputs, like air_model.py:479-493 does with its own KL
    g = torch.randn(B, 784)
    (orec * g).sum().add(0.5 * torch.sum(torch.exp(olv) + omean ** 2 - olv)).backward()
    ((rec * g.to(DEV)).sum() + 0.5 * torch.sum(torch.exp(logvar) + mean ** 2 - logvar)).backward()
    assert relnorm(xg.grad, xt.grad) < 1e-4


# ---------------------------------------------------------------------------------------------
# tcgen05 TF32 mode
# ---------------------------------------------------------------------------------------------
_WS = {}


def _gemm_ws():
    """caller-owned split-K scratch (air_gemm_ws): the library allocates nothing itself"""
    if "ws" not in _WS:
        _WS["ws"] = torch.zeros(16 << 20, device=DEV)   # zero once: its tail holds the split-K tickets
    return _WS["ws"]


def _tf32_case(M, N, Kd, tA, tB, seed=0, ints=True, epi=K.EPI_NONE, use_cinit=False, use_bias=False, mode="tf32"):
    rng = np.random.RandomState(seed)
    if ints:   # small integers are exact in TF32 and their dot products are exact in FP32
        A = rng.randint(-2, 3, (M, Kd)).astype(np.float32)
        Bm = rng.randint(-2, 3, (Kd, N)).astype(np.float32)
    else:
        A = rng.randn(M, Kd).astype(np.float32)
        Bm = rng.randn(Kd, N).astype(np.float32)
    lda = ((Kd if not tA else M) + 3) // 4 * 4
    ldb = ((N if not tB else Kd) + 3) // 4 * 4
    Abuf = torch.zeros((M if not tA else Kd), lda, device=DEV)
    Bbuf = torch.zeros((Kd if not tB else N), ldb, device=DEV)
    Av = Abuf[:, :(Kd if not tA else M)]
    Bv = Bbuf[:, :(N if not tB else Kd)]
    Av.copy_(cu(A.T.copy() if tA else A))
    Bv.copy_(cu(Bm.T.copy() if tB else Bm))
    Ci = cu(rng.randint(-3, 4, (M, N)).astype(np.float32)) if use_cinit else None
    bias = cu(rng.randint(-3, 4, N).astype(np.float32)) if use_bias else None
    out = torch.full((M, N), 7.0, device=DEV)
    ops.gemm(Av, Bv, out, Cinit=Ci, bias=bias, tA=tA, tB=tB, epi=epi, mode=K.GEMM_MODES[mode], ws=_gemm_ws())
    want = A.astype(np.float64) @ Bm.astype(np.float64)
    if use_cinit:
        want = want + Ci.cpu().numpy()
    if use_bias:
        want = want + bias.cpu().numpy()
    return out.cpu().numpy(), want


TC_MODES = ["tf32", "tf32x3"]


@pytest.mark.parametrize("mode", TC_MODES)
@pytest.mark.parametrize("tA,tB", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,Kd", [(128, 128, 32), (128, 64, 64), (256, 256, 256), (4096, 1024, 256), (200, 100, 50),
                                    (130, 70, 36), (784, 512, 4096), (2500, 1024, 512), (64, 8, 24), (4096, 320, 256)])
def test_gemm_tf32_exact_on_integers(M, N, Kd, tA, tB, mode):
    got, want = _tf32_case(M, N, Kd, tA, tB, seed=M + N + Kd, mode=mode)
    assert np.array_equal(got, want.astype(np.float32)), np.abs(got - want).max()


@pytest.mark.parametrize("mode", TC_MODES)
@pytest.mark.parametrize("tA,tB", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,Kd", [(4096, 640, 2112), (2500, 1024, 2500)])
def test_gemm_tf32_cluster_multicast_exact_on_integers(M, N, Kd, tA, tB, mode):
    """>= 64 k-blocks per CTA, even M-tile count, >= 148 tiles (no split-K): the 2-CTA cluster path, in which each CTA
    loads half of the shared B tile and TMA-multicasts it to both.  Ragged M (2500 = 19.5 tiles) and ragged K included."""
    got, want = _tf32_case(M, N, Kd, tA, tB, seed=M + N + Kd, use_cinit=True, use_bias=True, mode=mode)
    assert np.array_equal(got, want.astype(np.float32)), np.abs(got - want).max()


@pytest.mark.parametrize("mode", TC_MODES)
def test_gemm_tf32_epilogue_cinit_bias_and_splitk(mode):
    for (M, N, Kd, tA, tB) in ((784, 512, 4096, True, False), (4096, 512, 784, False, False), (256, 1024, 4096, True, False)):
        got, want = _tf32_case(M, N, Kd, tA, tB, seed=1, use_cinit=True, use_bias=True, epi=K.EPI_RELU, mode=mode)
        assert np.array_equal(got, np.maximum(want, 0).astype(np.float32))


@pytest.mark.parametrize("tA,tB", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,Kd", [(512, 384, 776), (4096, 1024, 2500), (784, 512, 12288), (256, 320, 12288), (130, 70, 36)])
def test_gemm_tf32x3_is_fp32_grade(M, N, Kd, tA, tB):
    """The 3xTF32 mode on random fp32 operands, every layout, the model's long-K shapes (CTA pairs, split-K): the same
    order of error against the fp64 product as the exact-FP32 FMA chain (measured 1-3e-6 vs 0.5-2e-6 norm-wise; what is
    left over FP32 is the round-toward-zero accumulate step of the tensor core, tests/diag_tc_rounding.py)."""
    got, want = _tf32_case(M, N, Kd, tA, tB, seed=5, ints=False, mode="tf32x3")
    ref32, _ = _tf32_case(M, N, Kd, tA, tB, seed=5, ints=False, mode="fp32")
    e3, e32 = relnorm(got, want), relnorm(ref32, want)
    assert e3 < 4e-6 and e3 < 4 * e32 + 1e-6, (e3, e32)
    # element-wise: no entry is off by more than 2e-5 of the typical magnitude sqrt(K)
    assert np.abs(got - want).max() < 2e-5 * np.sqrt(Kd)


def test_gemm_tf32x3_wide_dynamic_range_and_specials():
    """operands spanning 2^-40 .. 2^40 (the split must be exact per element, not per tile) and an inf / NaN entry
    (they travel in the hi part; the lo part of a non-finite value is 0 rather than inf - inf)"""
    rng = np.random.RandomState(7)
    M, N, Kd = 256, 128, 512
    A = (rng.randn(M, Kd) * np.exp2(rng.randint(-40, 41, (M, 1)))).astype(np.float32)
    Bm = (rng.randn(Kd, N) * np.exp2(rng.randint(-40, 41, (1, N)))).astype(np.float32)
    out = torch.empty(M, N, device=DEV)
    ops.gemm(cu(A), cu(Bm), out, mode=K.GEMM_MODES["tf32x3"])
    want = A.astype(np.float64) @ Bm.astype(np.float64)
    scale = np.sqrt((A.astype(np.float64) ** 2).sum(1))[:, None] * np.sqrt((Bm.astype(np.float64) ** 2).sum(0))[None, :]
    assert (np.abs(out.cpu().numpy() - want) / scale).max() < 1e-6
    A[3, 5] = np.inf
    A[7, 9] = np.nan
    ops.gemm(cu(A), cu(Bm), out, mode=K.GEMM_MODES["tf32x3"])
    got = out.cpu().numpy()
    assert (~np.isfinite(got[3])).all() and np.isnan(got[7]).all() and np.isfinite(np.delete(got, (3, 7), axis=0)).all()


def test_gemm_tf32_random_accuracy():
    """TF32 keeps 10 mantissa bits of each input: ~1e-3 relative on a dot product."""
    got, want = _tf32_case(512, 384, 777 // 4 * 4, False, False, seed=2, ints=False)
    assert relnorm(got, want) < 2e-3
    # inputs already rounded to TF32 -> only FP32 accumulation error remains
    rng = np.random.RandomState(3)
    A = torch.from_numpy(rng.randn(300, 256).astype(np.float32))
    Bm = torch.from_numpy(rng.randn(256, 200).astype(np.float32))
    trunc = lambda t: (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
    A, Bm = trunc(A), trunc(Bm)
    out = torch.empty(300, 200, device=DEV)
    ops.gemm(A.to(DEV), Bm.to(DEV), out, mode=K.GEMM_MODES["tf32"])
    assert relnorm(out, A.double() @ Bm.double()) < 1e-6


@pytest.mark.parametrize("mode", TC_MODES)
def test_gemm_tf32_rejects_unaligned_leading_dimension(mode):
    a = torch.zeros(64, 50, device=DEV)       # ld = 50 floats = 200 bytes: not a TMA stride
    with pytest.raises(ab.AirError, match="TMA|aligned"):
        ops.gemm(a, torch.zeros(50, 64, device=DEV), torch.zeros(64, 64, device=DEV), mode=K.GEMM_MODES[mode])


def test_gemm_split_k_needs_a_caller_workspace_and_is_deterministic():
    """without a workspace the long-K weight-gradient shape runs unsplit; with one it splits (a second pass sums the
    partials in split order); both are deterministic; a workspace too small for any split silently means 'unsplit'
    (never an allocation).  The split result is the more
    accurate one: the tensor core's accumulate step rounds toward zero, so an unsplit K = 12288 chain (3 x 128
    k-blocks) carries ~1e-5 of bias, the 32 short chains of the split run ~1e-6 (include/air_b200.h, AIR_GEMM_TF32X3)."""
    rng = np.random.RandomState(11)
    A, Bm = cu(rng.randn(12288, 256).astype(np.float32)), cu(rng.randn(12288, 320).astype(np.float32))
    outs = []
    for ws in (None, _gemm_ws(), _gemm_ws(), torch.empty(1000, device=DEV)):
        out = torch.empty(256, 320, device=DEV)
        n0 = ab.launch_count()
        ops.gemm(A, Bm, out, tA=True, mode=K.GEMM_MODES["tf32x3"], ws=ws)
        outs.append((out, ab.launch_count() - n0))
    assert [n for _, n in outs] == [1, 2, 2, 1]
    assert torch.equal(outs[1][0], outs[2][0]) and torch.equal(outs[0][0], outs[3][0]) and not torch.equal(outs[0][0], outs[1][0])
    want = A.double().t() @ Bm.double()
    e_unsplit, e_split = relnorm(outs[0][0], want), relnorm(outs[1][0], want)
    assert e_split < 2e-6 and e_unsplit < 2e-5 and e_split < e_unsplit, (e_split, e_unsplit)


# ---------------------------------------------------------------------------------------------
# counter-based noise generated in the kernels (csrc/rng.cuh) vs its host restatement (oracle/rng_oracle.py)
# ---------------------------------------------------------------------------------------------
def test_device_noise_generator_equals_host_restatement_bit_for_bit():
    from oracle import rng_oracle as R
    seed, L, TB = 0x1234567890ABCDEF & 0x7FFFFFFFFFFFFFFF, 50, 3 * 37
    st = ops.rng_state(seed, DEV, counter=41)
    # samples of a stream at the current counter, odd lengths included
    for stream, n in ((K.RNG_SHIFT, 4099), (K.RNG_LATENT, 777), (K.RNG_SCALE, 5)):
        assert np.array_equal(ops.rng_normals(st, stream, n).cpu().numpy(), R.normals(seed, 41, stream, n))
    # the likelihood stream (generated inside the gen_mean GEMM epilogue) uses the SFU form of the transform: same
    # Philox words, samples equal to the host formula to ~1e-5 (hardware lg2 / rsqrt / sin / cos approximations)
    like = ops.rng_normals(st, K.RNG_LIKE, 100_003).cpu().numpy()
    assert np.abs(like - R.normals(seed, 41, K.RNG_LIKE, 100_003)).max() < 3e-5
    assert np.array_equal(ops.rng_uniforms(st, K.RNG_CONCRETE, 1001).cpu().numpy(), R.uniforms(seed, 41, K.RNG_CONCRETE, 1001))
    # one fill = one step: the counter advances first, then every tensor is that counter's stream
    sc, sh, la, cu_ = (torch.empty(TB, device=DEV), torch.empty(TB, 2, device=DEV), torch.empty(TB, L, device=DEV),
                       torch.empty(TB, device=DEV))
    for step in (42, 43):
        ops.noise_fill(st, sc, sh, la, cu_, TB, L)
        assert st.cpu().tolist() == [seed, step, 0, 0]
        assert np.array_equal(sc.cpu().numpy(), R.normals(seed, step, K.RNG_SCALE, TB))
        assert np.array_equal(sh.cpu().numpy().ravel(), R.normals(seed, step, K.RNG_SHIFT, 2 * TB))
        assert np.array_equal(la.cpu().numpy().ravel(), R.normals(seed, step, K.RNG_LATENT, TB * L))
        assert np.array_equal(cu_.cpu().numpy(), R.uniforms(seed, step, K.RNG_CONCRETE, TB))
    # moments of a large draw (the transform is an exact Box-Muller up to fp32 rounding)
    z = ops.rng_normals(st, K.RNG_LIKE, 4_000_000).double()
    assert abs(z.mean()) < 2e-3 and abs(z.std() - 1) < 2e-3 and abs((z ** 4).mean() - 3) < 2e-2 and z.abs().max() < 6
    u = ops.rng_uniforms(st, K.RNG_CONCRETE, 1_000_000)
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 2e-3


@pytest.mark.parametrize("mode", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("M,N,Kd", [(300, 784, 512), (64, 50, 36), (256, 128, 0)])
def test_gemm_epilogue_generates_the_likelihood_noise(mode, M, N, Kd):
    """AIR_EPI_SIGMOID_RNG == AIR_EPI_SIGMOID_NOISE fed with the samples of stream RNG_LIKE (element m * N + n), in every
    GEMM path (tensor-core fast path, ragged tiles, SIMT, the K = 0 kernel): no noise tensor needed."""
    rng = np.random.RandomState(M + N)
    # (operands as views of padded buffers: TMA leading dimensions, and a K = 0 view keeps a unit inner stride)
    A = cu(rng.randn(M, (Kd + 3) // 4 * 4 + 4).astype(np.float32))[:, :Kd]
    Bm = cu(rng.randn(Kd + 1, (N + 3) // 4 * 4).astype(np.float32))[:Kd, :N]
    bias = cu(rng.randn(N).astype(np.float32))
    st = ops.rng_state(99, DEV, counter=7)
    noise = ops.rng_normals(st, K.RNG_LIKE, M * N).reshape(M, N)
    got, want = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
    ops.gemm(A, Bm, got, bias=bias, aux=st, epi=K.EPI_SIGMOID_RNG, epi_param=0.3, mode=K.GEMM_MODES[mode], ws=_gemm_ws())
    ops.gemm(A, Bm, want, bias=bias, aux=noise, epi=K.EPI_SIGMOID_NOISE, epi_param=0.3, mode=K.GEMM_MODES[mode], ws=_gemm_ws())
    assert torch.equal(got, want)

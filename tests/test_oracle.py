"""CPU tests pinning the oracle (no GPU): golden vectors, cross-restatement agreement,
known-answer cases, fp64 finite differences.  TensorFlow cannot run here; the pin against
the reference's own graph is tests/test_reference_graph.py."""
import os

import numpy as np
import pytest
import torch

from oracle import air_oracle as O
from oracle import c_oracle as C


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_golden_st(golden_dir):
    g = _g(golden_dir, "st.npz")
    for impl in ("torch", "c"):
        f = (lambda U, t, s: O.transformer(torch.from_numpy(U), torch.from_numpy(t), s).numpy()) if impl == "torch" \
            else C.st_forward
        assert np.array_equal(f(g["U"], g["theta"], (28, 28)), g["crop"])
        assert np.array_equal(f(g["window"], g["theta_inv"], (50, 50)), g["back"])
        assert np.array_equal(f(g["Ug"], g["thg"], (7, 9)), g["outg"])
    out = C.canvas_update(g["canvas"], g["back"].reshape(-1, 2500), g["z"], g["stop"], 0.99)
    assert np.array_equal(out, g["canvas_out"])
    # rows with stop >= thr are untouched (0.99 itself is NOT < 0.99)
    dead = g["stop"] >= np.float32(0.99)
    assert dead.sum() == 3 and np.array_equal(out[dead], g["canvas"][dead])


def test_golden_concrete(golden_dir):
    g = _g(golden_dir, "concrete.npz")
    for train in (0, 1):
        out = C.concrete_step(g["log_odds"], g["u"], g["stop_prev"], g["loss_prev"], g["digits_prev"],
                              float(g["prior"]), float(g["temperature"]), float(g["thr"]), train)
        for k, v in out.items():
            assert np.array_equal(v, g[f"{k}_train{train}"]), k
    # torch restatement agrees with the C one to float rounding of log/exp
    lo, u = torch.from_numpy(g["log_odds"]), torch.from_numpy(g["u"])
    y = O.concrete_binary_pre_sigmoid_sample(lo, 1.0, u)
    np.testing.assert_allclose(y.numpy(), g["y_train1"], rtol=2e-6, atol=2e-6)
    kl = O.concrete_binary_kl_mc_sample(y, float(g["prior"]), 1.0, lo, 1.0)
    np.testing.assert_allclose(kl.numpy(), g["kl_train1"], rtol=1e-5, atol=2e-6)


def test_golden_model(golden_dir):
    g = _g(golden_dir, "model_b8.npz")
    imgs, cnt = O.synthetic_canvases(8, seed=0)
    for train in (True, False):
        m = O.AIROracle(annealing_schedules=O.DEFAULT_ANNEALING, train=train, seed=0)
        m.global_step = 2000
        out = m.forward(imgs, cnt, O.make_noise(0, 3, 8))
        tag = "train" if train else "test"
        assert np.array_equal(out["rec_num_digits"].numpy(), g[f"{tag}_rec_num_digits"])
        assert np.array_equal(out["stop_masks"].numpy(), g[f"{tag}_stop_masks"])
        assert int(out["executed_steps"]) == int(g[f"{tag}_executed_steps"])
        # same machine + same BLAS gives bit equality; allow last-bit BLAS differences elsewhere
        np.testing.assert_allclose(out["loss"].numpy(), g[f"{tag}_loss"], rtol=1e-5)
        np.testing.assert_allclose(out["rec_scales"].numpy(), g[f"{tag}_rec_scales"], rtol=1e-5, atol=1e-6)


def test_identity_theta_is_1001_shrink():
    """Identity theta does NOT reproduce U: (W - 1.001) shrink, SURVEY 8a-2."""
    torch.manual_seed(0)
    U = torch.rand(2, 50, 50, 1)
    th = torch.tensor([[1.0, 0, 0, 0, 1.0, 0]]).repeat(2, 1)
    out = O.transformer(U, th, (50, 50))
    err = (out - U).abs().max().item()
    assert 1e-4 < err < 3e-3
    assert torch.equal(out[:, 0, 0], U[:, 0, 0])  # the (-1,-1) corner is exact


def test_out_of_range_residue_semantics():
    """Both axes clipped -> exactly 0; one axis clipped -> tiny rounding residue (hard part 1)."""
    U = torch.ones(1, 28, 28, 1)
    s = 0.5
    thi = torch.tensor([[1 / s, 0, 0.0, 0, 1 / s, 0.0]])
    out = O.transformer(U, thi, (50, 50))[0, :, :, 0]
    assert out[0, 0].item() == 0.0 and out[49, 49].item() == 0.0
    assert abs(out[25, 25].item() - 1.0) < 1e-5
    edge = out[25, 0].abs().item()  # x clipped, y inside
    assert edge < 1e-5


def test_crop_writeback_roundtrip():
    """writeback(crop(img)) reproduces a smooth image inside the attended window."""
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 50), torch.linspace(0, 1, 50), indexing="ij")
    img = (0.5 + 0.5 * torch.sin(3 * xx) * torch.cos(2 * yy)).reshape(1, 50, 50, 1)
    s, x, y = 0.6, 0.1, -0.2
    th = torch.tensor([[s, 0, x, 0, s, y]])
    thi = torch.tensor([[1 / s, 0, -x / s, 0, 1 / s, -y / s]])
    back = O.transformer(O.transformer(img, th, (28, 28)), thi, (50, 50))
    c = slice(18, 28)
    assert (back[0, c, 22:32, 0] - img[0, c, 22:32, 0]).abs().max() < 0.02


def test_st_gradcheck_fp64():
    torch.manual_seed(1)
    U = torch.rand(2, 9, 8, 2, dtype=torch.float64, requires_grad=True)
    th = (torch.tensor([[0.7, 0.1, 0.05, -0.1, 0.8, -0.03]], dtype=torch.float64).repeat(2, 1)
          + 0.01 * torch.randn(2, 6, dtype=torch.float64)).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda a, b: O.transformer(a, b, (5, 6)), (U, th), eps=1e-7, atol=1e-5)


def test_c_backward_matches_autograd():
    rng = np.random.RandomState(0)
    U = rng.rand(3, 12, 10, 2).astype(np.float32)
    th = (np.array([[0.8, 0.1, 0.0, -0.1, 0.7, 0.1]], np.float32) + 0.05 * rng.randn(3, 6)).astype(np.float32)
    g = rng.randn(3, 6, 7, 2).astype(np.float32)
    Ut, tt = torch.from_numpy(U).requires_grad_(True), torch.from_numpy(th).requires_grad_(True)
    O.transformer(Ut, tt, (6, 7)).backward(torch.from_numpy(g))
    dU, dth = C.st_backward(U, th, g)
    np.testing.assert_allclose(dU, Ut.grad.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dth.reshape(3, 6), tt.grad.numpy(), rtol=1e-4, atol=1e-4)


def test_concrete_kl_closed_form():
    """log-density restatement vs an fp64 closed form of the binary Concrete density."""
    y = torch.linspace(-4, 4, 33, dtype=torch.float64)
    a_q, a_p, tau = 0.7, -2.0, 1.3

    def logpdf(y, a, t):
        return np.log(t) - t * y + a - 2 * np.log1p(np.exp(-t * y + a))

    want = logpdf(y.numpy(), a_q, tau) - logpdf(y.numpy(), a_p, tau)
    got = O.concrete_binary_kl_mc_sample(y, a_p, tau, torch.tensor(a_q, dtype=torch.float64), tau)
    np.testing.assert_allclose(got.numpy(), want, rtol=1e-6, atol=1e-6)


def test_stop_semantics_zero_digits():
    """z ~ 0 at step 1 => stop ~ 1 >= 0.99 => zero digits (README: thr = 0.99 < 1)."""
    out = C.concrete_step(np.array([-30.0, 30.0], np.float32), np.array([0.5, 0.5], np.float32),
                          np.zeros(2, np.float32), np.zeros(2, np.float32), np.zeros(2, np.int32), -0.01, 1.0, 0.99, 1)
    assert out["digits_new"].tolist() == [0, 1]
    assert out["stop_new"][0] >= 0.99 and out["stop_new"][1] < 0.01


def test_round_half_even_at_test_time():
    out = C.concrete_step(np.array([0.0], np.float32), np.array([0.5], np.float32), np.zeros(1, np.float32),
                          np.zeros(1, np.float32), np.zeros(1, np.int32), 0.0, 1.0, 0.99, 0)
    assert out["z"][0] == 0.0  # sigmoid(0) = 0.5 rounds to even (0), tf.round semantics


def test_fixed_trip_equals_early_exit():
    """Items that stopped contribute exact +0.0: digits cannot grow after stopping."""
    imgs, cnt = O.synthetic_canvases(16, seed=5)
    m = O.AIROracle(train=False, seed=1, max_steps=5)
    out = m.forward(imgs, cnt, O.make_noise(1, 5, 16))
    masks = out["stop_masks"].numpy()
    assert np.all(np.diff(masks.astype(np.int32), axis=1) <= 0)  # live mask is monotone non-increasing
    assert np.array_equal(out["rec_num_digits"].numpy(), masks.sum(1))


def test_tf_softplus_branches_and_grad():
    x = torch.tensor([-20.0, -13.0, 0.0, 13.0, 20.0], requires_grad=True)
    y = O.tf_softplus(x)
    assert y[0].item() == pytest.approx(np.exp(-20.0), rel=1e-6)
    assert y[4].item() == 20.0
    y.sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), 1 / (1 + np.exp(-x.detach().numpy())), rtol=1e-6)


def test_tf_adam_and_clip_semantics():
    m = O.AIROracle(seed=0)
    g = {k: torch.full_like(v, 0.5) for k, v in m.params.items()}
    clipped, norm = O.AIROracle.clip_by_global_norm(g, 1.0)
    n = sum(v.numel() for v in g.values())
    assert norm.item() == pytest.approx(0.5 * np.sqrt(n), rel=1e-5)
    k0 = "rnn/bias"
    before = m.params[k0].clone()
    m.adam_apply(clipped)
    gc = 0.5 / norm.item()
    mm, vv = 0.1 * gc, 0.001 * gc * gc
    alpha = 1e-4 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = -alpha * mm / (np.sqrt(vv) + 1e-8)   # epsilon OUTSIDE the bias correction
    np.testing.assert_allclose((m.params[k0] - before).numpy(), want, rtol=1e-4)
    assert m.global_step == 1


def test_annealed_prior():
    v = O.annealed_value(O.DEFAULT_ANNEALING["z_pres_prior_log_odds"], 3000)
    assert v.item() == pytest.approx(np.log(1000.0), rel=1e-5)
    v = O.annealed_value(O.DEFAULT_ANNEALING["z_pres_prior_log_odds"], 60000)
    assert v.item() == pytest.approx(np.log(2e-9), rel=1e-4)   # min 1e-9, then log(. + 1e-9)


def test_param_census_matches_checkpoint_index():
    """36 trainables, 4,011,643 parameters (model/air-model.index)."""
    shapes = O.param_shapes()
    assert len(shapes) == 36
    assert sum(int(np.prod(s)) for s in shapes.values()) == 4011643
    assert shapes["rnn/kernel"] == (2756, 1024)


def test_gemm_seq_fma_oracle():
    rng = np.random.RandomState(0)
    A, Bm = rng.randn(5, 37).astype(np.float32), rng.randn(37, 11).astype(np.float32)
    out = C.gemm_seq_fma(A, Bm, bias=np.ones(11, np.float32))
    np.testing.assert_allclose(out, A.astype(np.float64) @ Bm.astype(np.float64) + 1, rtol=1e-5, atol=1e-5)


# ---- hypothesis: the two independent restatements agree bit for bit on arbitrary shapes / thetas ----------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=40, deadline=None)
@given(B=st.integers(1, 4), H=st.integers(2, 17), W=st.integers(2, 19), Cc=st.integers(1, 3), oh=st.integers(1, 12),
       ow=st.integers(1, 13), seed=st.integers(0, 10 ** 6), scale=st.floats(0.05, 3.0), axis_aligned=st.booleans())
def test_c_and_torch_oracles_agree_bitwise(B, H, W, Cc, oh, ow, seed, scale, axis_aligned):
    rng = np.random.RandomState(seed)
    U = rng.rand(B, H, W, Cc).astype(np.float32)
    th = (rng.uniform(-1, 1, (B, 2, 3)) * scale).astype(np.float32)
    if axis_aligned:
        th[:, 0, 1] = 0.0
        th[:, 1, 0] = 0.0
    a = C.st_forward(U, th, (oh, ow))
    b = O.transformer(torch.from_numpy(U), torch.from_numpy(th), (oh, ow)).numpy()
    assert np.array_equal(a, b)
    assert np.isfinite(a).all()
    # out-of-range samples never exceed rounding-residue magnitude times the image scale
    assert np.abs(a).max() <= U.max() * (1 + 1e-5) + 1e-5


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 40), seed=st.integers(0, 10 ** 6), train=st.booleans(), tau=st.floats(0.3, 2.0))
def test_concrete_step_invariants(n, seed, train, tau):
    rng = np.random.RandomState(seed)
    lo = (rng.randn(n) * 3).astype(np.float32)
    u = rng.rand(n).astype(np.float32)
    sp = rng.choice(np.array([0.0, 0.4, 0.98, 0.99, 2.0], np.float32), n)
    out = C.concrete_step(lo, u, sp, np.zeros(n, np.float32), np.zeros(n, np.int32), -0.5, tau, 0.99, int(train))
    assert np.all((out["z"] >= 0) & (out["z"] <= 1))
    if not train:
        assert np.all((out["z"] == 0) | (out["z"] == 1))
    assert np.all(out["stop_new"] >= sp)                                       # the stopping sum never decreases
    assert np.array_equal(out["digits_new"], (out["stop_new"] < np.float32(0.99)).astype(np.int32))
    assert np.array_equal(out["loss_new"] != 0, (sp < np.float32(0.99)) & (out["kl"] != 0))   # KL masked by the OLD sum


def test_cnn_frontend_shapes_and_store_init_match_oracle():
    """cnn=True (air_model.py:510-535): 50x50x1 -> conv/pool 25x25x8 -> conv/pool 12x12x8 -> conv 12x12x8 -> [B,1152];
    the host ParamStore initialises the same tensors, in the same RNG order, as the oracle."""
    import air_b200 as ab
    from air_b200 import checkpoint
    p = O.init_params(seed=4, cnn=True)
    assert tuple(p["cnn/conv1/kernel"].shape) == (5, 5, 1, 8) and tuple(p["rnn/kernel"].shape) == (1152 + 256, 1024)
    x = torch.rand(3, 2500)
    f = O.cnn_frontend(x, p)
    assert tuple(f.shape) == (3, 1152) and float(f.min()) >= 0.0
    # NHWC flatten: feature (y, x, c) of image b sits at index (y*12 + x)*8 + c
    import torch.nn.functional as F
    t = torch.relu(F.conv2d(x.reshape(3, 1, 50, 50), p["cnn/conv1/kernel"].permute(3, 2, 0, 1), p["cnn/conv1/bias"], padding=2))
    t = F.max_pool2d(t, 2, 2)
    t = F.max_pool2d(torch.relu(F.conv2d(t, p["cnn/conv2/kernel"].permute(3, 2, 0, 1), p["cnn/conv2/bias"], padding=2)), 2, 2)
    t = torch.relu(F.conv2d(t, p["cnn/conv3/kernel"].permute(3, 2, 0, 1), p["cnn/conv3/bias"], padding=2))
    assert torch.equal(f.reshape(3, 12, 12, 8)[1, 5, 7, 3], t[1, 3, 5, 7])
    store = ab.ParamStore("cpu", 1152, 784, 256, 64, 50, (512, 256), (256, 512), seed=4, cnn_filters=8)
    named = store.named_views()
    assert list(named)[:6] == [f"cnn/conv{i}/{n}" for i in (1, 2, 3) for n in ("kernel", "bias")]
    for k, v in p.items():
        assert torch.equal(named[k], v), k
    names = checkpoint.model_tensors(store)
    assert "air/cnn/conv1/kernel" in names and "air/rnn/rnn/kernel" in names and "air/training/air/cnn/conv2/bias/Adam" in names
    # an oracle forward pass with the front-end runs end to end and differs from the plain model
    imgs, cnt = O.synthetic_canvases(4, seed=0)
    noise = O.make_noise(1, 3, 4)
    out = O.AIROracle(params=p, cnn=True).forward(imgs, cnt, noise)
    assert torch.isfinite(out["loss"])


def test_synth_oracle_properties():
    """The restated device canvas generator: digit counts uniform over {0,1,2}, exact-zero background, no overlap
    clash (a canvas with k digits has k connected blobs' worth of pixels), image i independent of the batch."""
    from oracle import synth_oracle as S
    im, cnt = S.synth_canvases(600, seed=2)
    assert im.shape == (600, 2500) and im.dtype == np.float32 and cnt.dtype == np.int32
    frac = np.bincount(cnt, minlength=3) / 600.0
    assert np.abs(frac - 1 / 3).max() < 0.07
    assert im.min() == 0.0 and 0.9 < im.max() <= 1.0
    assert not im[cnt == 0].any() and (im[cnt > 0] > 0).any(1).all()
    a, ca = S.synth_canvases(10, seed=2, first_index=100)
    assert np.array_equal(a, im[100:110]) and np.array_equal(ca, cnt[100:110])
    more = (im[cnt == 2] > 0).sum(1).mean() > 1.5 * (im[cnt == 1] > 0).sum(1).mean()
    assert more                                                         # two digits -> about twice the ink


def test_cnn_frontend_golden(golden_dir):
    """oracle.cnn_frontend against the committed vector (tests/golden/make_golden_cnn.py): features and the
    gradients of all six conv tensors."""
    g = np.load(os.path.join(golden_dir, "cnn.npz"))
    leaf = {k[2:]: torch.from_numpy(g[k]).requires_grad_(True) for k in g.files if k.startswith("p:")}
    feat = O.cnn_frontend(torch.from_numpy(g["x"]), leaf)
    assert tuple(feat.shape) == (3, 1152) and float((feat > 0).float().mean()) > 0.1
    np.testing.assert_allclose(feat.detach().numpy(), g["features"], rtol=1e-5, atol=1e-6)
    (feat * torch.from_numpy(g["G"])).sum().backward()
    for k, v in leaf.items():
        np.testing.assert_allclose(v.grad.numpy(), g["g:" + k], rtol=1e-4, atol=1e-5, err_msg=k)



def test_cnn_c_restatement_agrees_with_torch_restatement(golden_dir):
    """Second, independent restatement of the cnn=True front-end (oracle/st_oracle.c::oracle_conv5x5_relu_pool*,
    plain loops) against oracle.cnn_frontend (torch conv2d / autograd) on the committed vector: the three layers
    chained forward, and the chained backward down to every kernel / bias gradient."""
    g = np.load(os.path.join(golden_dir, "cnn.npz"))
    P = {k[2:]: g[k] for k in g.files if k.startswith("p:")}
    layers = [("cnn/conv1", True), ("cnn/conv2", True), ("cnn/conv3", False)]
    acts = [g["x"].reshape(3, 50, 50, 1)]
    for name, pool in layers:
        acts.append(C.conv5x5_relu_pool(acts[-1], P[name + "/kernel"], P[name + "/bias"], pool))
    assert acts[1].shape == (3, 25, 25, 8) and acts[2].shape == (3, 12, 12, 8) and acts[3].shape == (3, 12, 12, 8)
    np.testing.assert_allclose(acts[3].reshape(3, 1152), g["features"], rtol=1e-5, atol=1e-6)
    d = g["G"].reshape(3, 12, 12, 8)
    for li in (2, 1, 0):
        name, pool = layers[li]
        d, dw, db = C.conv5x5_relu_pool_bwd(acts[li], P[name + "/kernel"], P[name + "/bias"], d, pool, need_dx=li > 0)
        np.testing.assert_allclose(dw, g["g:" + name + "/kernel"], rtol=1e-4, atol=1e-5, err_msg=name)
        np.testing.assert_allclose(db, g["g:" + name + "/bias"], rtol=1e-4, atol=1e-5, err_msg=name)
    # 'same' padding known answer: an all-ones 5x5x1x1 kernel on an all-ones 4x4 image counts the in-range taps
    ones = C.conv5x5_relu_pool(np.ones((1, 4, 4, 1), np.float32), np.ones((5, 5, 1, 1), np.float32),
                               np.zeros(1, np.float32), False)[0, :, :, 0]
    assert np.array_equal(ones, np.outer([3, 4, 4, 3], [3, 4, 4, 3]).astype(np.float32))


def test_synth_oracle_equals_the_references_generate_multi_image(golden_dir):
    """tests/golden/ref_multi_mnist_gen.npz holds what the reference's OWN generate_multi_image (multi_mnist.py:82-183)
    returned for the blobs and placement draws of images 5..100 of seed 11 (make_golden_multi_mnist.py): canvases bit
    for bit, digit counts, positions (x, y) and boxes (w, h).  The restated generator must reproduce them exactly --
    and the device generator is bit-exact with the restatement (tests/test_gpu_ops.py)."""
    from oracle import synth_oracle as S
    g = np.load(os.path.join(golden_dir, "ref_multi_mnist_gen.npz"))
    n = len(g["counts"])
    im, cnt, pos, box = S.synth_canvases(n, seed=int(g["seed"]), first_index=int(g["first_index"]), with_boxes=True)
    assert np.array_equal(cnt, g["counts"]) and np.array_equal(im, g["canvases"])
    assert np.array_equal(pos, g["positions"]) and np.array_equal(box, g["boxes"])
    assert set(np.unique(cnt)) == {0, 1, 2} and (box[cnt == 2] > 0).all()


@pytest.mark.skipif(not os.path.exists("/root/reference/multi_mnist.py"), reason="/root/reference not present (GPU box)")
def test_references_generator_replayed_live_including_a_canvas_restart():
    """The same replay, live, on other images -- and on a crowded 36x36 canvas with up to three digits, where placements
    fail after 100 attempts and the reference starts the canvas over (multi_mnist.py:95-171): same restarts, same
    result."""
    from oracle import synth_oracle as S
    from tests.golden import make_golden_multi_mnist as G
    ref = G.load_reference()
    for img in (1000, 1001, 1002, 1003, 1004, 1005):
        canvas, count, _, pos, box = G.reference_canvas(ref, 3, img)
        c2, k2, p2, b2 = S.one_canvas(3, img)
        assert count == k2 and np.array_equal(canvas, c2) and list(pos) == list(p2[:2 * k2]) and list(box) == list(b2[:2 * k2])
    restarted = 0
    for img in range(40):
        trace = {}
        c2, k2, p2, b2 = S.one_canvas(7, img, cs=36, max_digits=3, trace=trace)
        restarted += len(trace["blobs"]) > k2
        canvas, count, _, pos, box = G.reference_canvas(ref, 7, img, cs=36, max_digits=3)
        assert count == k2 and np.array_equal(canvas, c2) and list(pos) == list(p2[:2 * k2]) and list(box) == list(b2[:2 * k2])
    assert restarted > 0, "the crowded configuration should exercise the restart path"


def test_rng_oracle_philox_known_answers_and_moments():
    """oracle/rng_oracle.py (the host restatement of csrc/rng.cuh): Philox4x32-10 against the known-answer vectors of
    the Random123 distribution (kat_vectors: zero, all-ones and pi-digits counter / key), and the Box-Muller transform's
    moments / independence of the two samples of a pair."""
    from oracle import rng_oracle as R
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        assert tuple(int(v) for v in R.philox4x32(*ctr, *key)) == want
    z = R.normals(5, 11, R.STREAM_LATENT, 2_000_000).astype(np.float64)
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 3e-3 and abs((z ** 3).mean()) < 1e-2 and abs((z ** 4).mean() - 3) < 3e-2
    assert abs(np.corrcoef(z[0::2], z[1::2])[0, 1]) < 5e-3
    u = R.uniforms(5, 11, R.STREAM_CONCRETE, 1_000_000)
    assert u.dtype == np.float32 and 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 2e-3
    # streams, counters and seeds decorrelate; the same arguments reproduce
    a = R.normals(5, 11, R.STREAM_LIKE, 1000)
    assert np.array_equal(a, R.normals(5, 11, R.STREAM_LIKE, 1000))
    for other in (R.normals(5, 12, R.STREAM_LIKE, 1000), R.normals(6, 11, R.STREAM_LIKE, 1000), R.normals(5, 11, R.STREAM_SCALE, 1000)):
        assert abs(np.corrcoef(a, other)[0, 1]) < 0.15
    assert np.array_equal(R.normals(5, 11, R.STREAM_LIKE, 1000)[:997], R.normals(5, 11, R.STREAM_LIKE, 997))

"""The reference's checked-in SOURCE, executed (no GPU, no TensorFlow): air/transformer.py and air/concrete.py are
imported from /root/reference with ``tensorflow`` resolved to oracle/tfgraph/tf_shim.py, a numpy stand-in for the ~30
``tf.*`` calls they make.  This pins the functions the saved graph does not contain -- batch_transformer
(transformer.py:178-195), concrete_binary_sample incl. hard=True (concrete.py:4-17) -- and re-pins transformer() for
multi-channel, non-square and rotated cases.  Skipped where /root/reference does not exist (the GPU box); the GPU tests
compare the CUDA path with the same oracle functions (tests/test_gpu_st.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import air_oracle as O
from oracle import c_oracle as C
from oracle.tfgraph import tf_shim as S

REF = "/root/reference/air/"
pytestmark = pytest.mark.skipif(not os.path.exists(REF + "transformer.py"), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    with S.installed():
        return S.load_reference_module(REF + "transformer.py"), S.load_reference_module(REF + "concrete.py")


def _thetas(rng, B):
    th = np.zeros((B, 6), np.float32)
    th[:, 0] = th[:, 4] = rng.uniform(0.05, 1.6, B)      # incl. windows larger than the canvas
    th[:, 2], th[:, 5] = rng.uniform(-0.6, 0.6, B), rng.uniform(-0.6, 0.6, B)
    th[::3] = rng.uniform(-1.1, 1.1, (len(th[::3]), 6))  # every third: shear / rotation
    return th


@pytest.mark.parametrize("shape,out", [((9, 50, 50, 1), (28, 28)), ((9, 28, 28, 1), (50, 50)), ((5, 20, 30, 3), (7, 9)),
                                       ((2, 1, 1, 1), (3, 3)), ((4, 50, 50, 1), (100, 100))])
def test_transformer_source_bit_exact(ref, shape, out):
    rng = np.random.default_rng(sum(shape))
    U = rng.uniform(0, 1, shape).astype(np.float32)
    th = _thetas(rng, shape[0])
    want = ref[0].transformer(U, th, out)
    assert want.shape == (shape[0],) + out + (shape[3],) and want.dtype == np.float32
    assert np.array_equal(C.st_forward(U, th, out), want)
    assert np.array_equal(O.transformer(torch.from_numpy(U), torch.from_numpy(th), out).numpy(), want)
    # theta as [B,2,3] is accepted (transformer.py:144 reshapes)
    assert np.array_equal(ref[0].transformer(U, th.reshape(-1, 2, 3), out), want)


def test_batch_transformer_source_bit_exact(ref):
    rng = np.random.default_rng(5)
    U = rng.uniform(0, 1, (3, 20, 30, 2)).astype(np.float32)
    ths = rng.uniform(-1, 1, (3, 4, 6)).astype(np.float32)
    want = ref[0].batch_transformer(U, S.tensor(ths), (7, 9))
    assert want.shape == (12, 7, 9, 2)
    got = O.batch_transformer(torch.from_numpy(U), torch.from_numpy(ths), (7, 9)).numpy()
    assert np.array_equal(got, want)
    # output row b*N + n is input b under its n-th transform
    one = ref[0].transformer(U[1:2], ths[1, 2:3], (7, 9))
    assert np.array_equal(want[1 * 4 + 2], one[0])


def test_concrete_source(ref):
    rng = np.random.default_rng(6)
    lo = rng.normal(0, 3, 256).astype(np.float32)
    u = rng.uniform(0, 1, 256).astype(np.float32)
    u[:4] = [0.0, 1.0 - 2.0 ** -24, 0.5, 1e-12]           # the eps terms matter at the ends of [0,1)
    t = np.float32(0.7)
    with S.installed(uniform=u):
        y = ref[1].concrete_binary_pre_sigmoid_sample(lo, t)
        y_soft, s_soft = ref[1].concrete_binary_sample(lo, t, hard=False)
        y_hard, s_hard = ref[1].concrete_binary_sample(lo, t, hard=True)
    kl = ref[1].concrete_binary_kl_mc_sample(y, np.float32(-1.3), t, lo, t)
    tl, tu = torch.from_numpy(lo), torch.from_numpy(u)
    oy = O.concrete_binary_pre_sigmoid_sample(tl, 0.7, tu)
    np.testing.assert_allclose(oy.numpy(), y, rtol=2e-6, atol=2e-6)          # numpy vs torch log differ by an ulp
    okl = O.concrete_binary_kl_mc_sample(torch.from_numpy(y), -1.3, 0.7, tl, 0.7)
    np.testing.assert_allclose(okl.numpy(), kl, rtol=1e-5, atol=1e-5)
    oy2, os2 = O.concrete_binary_sample(tl, 0.7, tu, hard=False)
    np.testing.assert_allclose(oy2.numpy(), y_soft, rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(os2.numpy(), s_soft, rtol=2e-6, atol=1e-7)
    oy3, os3 = O.concrete_binary_sample(tl, 0.7, tu, hard=True)
    safe = np.abs(y_hard / t) > 1e-4                                         # away from sigmoid == 0.5 exactly
    assert np.array_equal(os3.numpy()[safe], s_hard[safe]) and set(np.unique(s_hard)) <= {0.0, 1.0}
    # pre-sigmoid sample and (y, sigmoid) variant are the same draw: y_soft / t == y
    np.testing.assert_allclose(y_soft / t, y, rtol=1e-6, atol=1e-6)
    # the C restatement of the fused step uses the same sample and KL
    out = C.concrete_step(lo, u, np.zeros(256, np.float32), np.zeros(256, np.float32), np.zeros(256, np.int32),
                          -1.3, 0.7, 0.99, 1)
    np.testing.assert_allclose(out["kl"], ref[1].concrete_binary_kl_mc_sample(y, np.float32(-1.3), t, lo, t),
                               rtol=1e-5, atol=1e-5)


def test_vae_source(ref):
    """air/vae.py through the shim (contrib.layers.fully_connected = act(x @ W + b), TF softplus, injected normal
    noise): reconstruction / mean / log-variance agree with the oracle's vae(), and the FOURTH return value of the
    checked-in source is the recognition MEAN (vae.py:43) -- the saved graph, an older revision, recorded the sample
    there; the oracle and the CUDA path follow the source."""
    rng = np.random.default_rng(7)
    params = O.init_params(seed=7)
    vp = {k[len("vae/"):]: v.numpy() for k, v in params.items() if k.startswith("vae/")}
    x = rng.uniform(0, 1, (16, 784)).astype(np.float32)
    n1, n2 = rng.standard_normal((16, 50)).astype(np.float32), rng.standard_normal((16, 784)).astype(np.float32)
    with S.installed(normal=[n1, n2], params=vp):
        vae_mod = S.load_reference_module(REF + "vae.py")
        rec, mean, logvar, latent = vae_mod.vae(x, 784, (512, 256), 50, (256, 512), 0.3)
    orec, omean, ologvar, olatent = O.vae(torch.from_numpy(x), params, "vae/", 2, 2, torch.from_numpy(n1),
                                          torch.from_numpy(n2), 0.3)
    assert latent is mean
    for got, want in ((orec, rec), (omean, mean), (ologvar, logvar), (olatent, mean)):
        np.testing.assert_allclose(got.numpy(), want, rtol=2e-5, atol=2e-6)


def _run_source_model(imgs, cnt, params, noise, T=3, cnn=False, global_step=0, train=False,
                      annealing=O.DEFAULT_ANNEALING, **kw):
    """AIRModel of the checked-in air/air_model.py (train=False: the optimizer is not built), executed eagerly by the
    shim: tf.while_loop is a Python loop, TensorArrays are lists, variables come from ``params``, the five noise
    tensors of every step are served in call order (scale, shift, VAE latent, VAE likelihood normals; Concrete uniform)."""
    import importlib
    import sys
    normal, uniform = [], []
    for t in range(T):
        normal += [noise[k][t].numpy() for k in ("scale", "shift", "vae_latent", "vae_like")]
        uniform.append(noise["concrete_u"][t].numpy())
    P = {k: v.numpy() for k, v in params.items()}
    with S.installed(uniform=uniform, normal=normal, params=P, global_step=global_step):
        sys.path.insert(0, "/root/reference")
        try:
            mod = importlib.import_module("air.air_model")
            hyper = dict(O.DEFAULT_HYPER, **kw)
            hyper.update(cnn=cnn, max_steps=T)
            m = mod.AIRModel(S.tensor(imgs.numpy()), S.tensor(cnt.numpy().astype(np.int32)), train=train,
                             annealing_schedules=annealing, **hyper)
            return m, dict(S.summaries)
        finally:
            sys.path.remove("/root/reference")
            for k in [k for k in sys.modules if k == "air" or k.startswith("air.")]:
                del sys.modules[k]


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


PER_STEP = ("rec_scales", "rec_shifts", "rec_st_back", "rec_windows", "rec_latents", "z_pres_probs", "z_pres_kls",
            "scale_kls", "shift_kls", "vae_kls")


@pytest.mark.parametrize("fixture,seed,T", [("realistic", 2, 3), ("default", 1, 3), ("realistic", 8, 5)])
def test_air_model_source_forward(fixture, seed, T):
    """The whole forward model of the CHECKED-IN source (air_model.py:269-611, test mode) against the oracle: digit
    counts exact, every per-step output <= 1e-5 -- including rec_latents, which at this revision is the recognition
    mean (the saved graph is older there) -- plus the 88 scalar summaries and the image summary the source builds."""
    import types
    import air_b200 as ab
    from tests import parity_util as PU
    imgs, cnt, params, noise = getattr(PU, fixture + "_fixture")(64, seed=seed, T=T)
    m, summ = _run_source_model(imgs, cnt, params, noise, T=T)
    out = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False, max_steps=T).forward(imgs, cnt, noise)
    assert np.array_equal(m.rec_num_digits, out["rec_num_digits"].numpy()) and len(np.unique(m.rec_num_digits)) >= 3
    for k in PER_STEP:
        got = np.asarray(getattr(m, k))
        assert got.shape == tuple(out[k].shape) and _rel(got, out[k].numpy()) < 1e-5, (k, _rel(got, out[k].numpy()))
    assert float(m.accuracy) == pytest.approx(float(out["accuracy"]), abs=1e-7)
    if T != 3:
        return                                            # the source's step summaries index max_steps columns
    # summaries: host logic of demo/model_wrapper.py and demo/visualize.py fed with the SOURCE model's own tensors
    stand = types.SimpleNamespace(max_digits=2, max_steps=3, **{
        k: torch.from_numpy(np.asarray(getattr(m, k))) for k in PER_STEP + ("rec_num_digits", "reconstruction_loss")})
    stand.loss_per_item = stand.reconstruction_loss + out["running_loss"]
    got = ab.evaluation_summaries(stand, cnt)
    scalars = {k: v for k, v in summ.items() if not k.startswith("image:")}
    assert set(got) == set(scalars) and len(scalars) == 88
    for k, v in scalars.items():
        tol = 1e-5 if "total_loss" not in k else 1e-4     # total_loss adds the oracle's running KL sum
        assert (np.isnan(v) and np.isnan(got[k])) or got[k] == pytest.approx(v, rel=tol, abs=1e-6), (k, got[k], v)
    vis = ab.visualize_reconstructions(imgs[:60], torch.from_numpy(np.asarray(m.reconstruction))[:60],
                                       torch.from_numpy(np.asarray(m.rec_st_back))[:60],
                                       torch.from_numpy(np.asarray(m.rec_num_digits))[:60], transformer=O.transformer)
    assert np.array_equal(vis.numpy(), summ["image:reconstruction"])


def test_air_model_source_covered_loss_and_cnn_frontend():
    from tests import parity_util as PU
    imgs, cnt, params, noise = PU.covered_fixture(64, seed=3)
    m, _ = _run_source_model(imgs, cnt, params, noise, global_step=2000)
    orc = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False)
    orc.global_step = 2000
    out = orc.forward(imgs, cnt, noise)
    assert abs(float(m.loss) - float(out["loss"])) <= 1e-6 * abs(float(out["loss"]))
    assert _rel(np.asarray(m.reconstruction_loss), out["reconstruction_loss"].numpy()) < 1e-6
    # cnn=True (air_model.py:510-535): the source's wiring of the front-end (three 5x5 convs, two 2x2 pools, NHWC
    # flatten to 12*12*8) with the shim's plain-numpy conv; features and the model on top of them follow the oracle
    imgs, cnt, params, noise = PU.covered_fixture(16, seed=4, cnn=True)
    m, _ = _run_source_model(imgs, cnt, params, noise, cnn=True)
    feat = O.cnn_frontend(imgs, params)
    assert np.asarray(m.rnn_input).shape == (16, 1152) and _rel(np.asarray(m.rnn_input), feat.numpy()) < 1e-5
    out = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False, cnn=True).forward(imgs, cnt, noise)
    assert np.array_equal(m.rec_num_digits, out["rec_num_digits"].numpy())
    for k in PER_STEP:
        assert _rel(np.asarray(getattr(m, k)), out[k].numpy()) < 1e-5, k


def test_air_model_source_train_mode_forward(golden_dir):
    """train=True forward of the source (continuous z_pres; the optimizer is a no-op stand-in, there is no autodiff in
    the shim): on the covered fixture the loss equals the oracle's AND the saved graph's (5730.3291), i.e. source,
    graph and oracle agree on the training objective; on realistic poses the per-step quantities and masks agree."""
    from tests import parity_util as PU
    g = np.load(os.path.join(golden_dir, "ref_graph_train_covered.npz"))
    imgs, cnt, params, noise = PU.covered_fixture(64, seed=3)
    m, _ = _run_source_model(imgs, cnt, params, noise, global_step=2000, train=True)
    assert m.training is None and abs(float(m.loss) - float(g["loss"])) <= 1e-6 * float(g["loss"])
    for k in ("rec_scales", "rec_shifts", "z_pres_kls", "scale_kls", "shift_kls", "vae_kls"):
        assert _rel(np.asarray(getattr(m, k)).reshape(g[k].shape), g[k]) < 1e-6, k
    imgs, cnt, params, noise = PU.realistic_fixture(64, seed=1)
    m, _ = _run_source_model(imgs, cnt, params, noise, global_step=2000, train=True)
    orc = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=True)
    orc.global_step = 2000
    out = orc.forward(imgs, cnt, noise)
    assert np.array_equal(m.rec_num_digits, out["rec_num_digits"].numpy()) and len(np.unique(m.rec_num_digits)) >= 3
    for k in PER_STEP:
        assert _rel(np.asarray(getattr(m, k)), out[k].numpy()) < 1e-5, k


def test_transformer_source_fuzz(ref):
    """Property test: for random shapes (1..6 images, 1..12 x 1..12 pixels, 1..3 channels, 1..9 x 1..9 outputs) and
    random affine maps -- tiny / huge scales, far-off shifts, exact identity, all-zero theta -- the C restatement and the
    torch restatement reproduce the source's transformer() bit for bit."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.integers(1, 6), st.integers(1, 12), st.integers(1, 12), st.integers(1, 3), st.integers(1, 9),
           st.integers(1, 9), st.integers(0, 2 ** 31 - 1), st.sampled_from(["affine", "identity", "zero", "huge", "tiny"]))
    def check(B, H, W, C_, oh, ow, seed, kind):
        rng = np.random.default_rng(seed)
        U = rng.uniform(-1, 1, (B, H, W, C_)).astype(np.float32)
        th = rng.uniform(-1.5, 1.5, (B, 6)).astype(np.float32)
        if kind == "identity":
            th[:] = [1, 0, 0, 0, 1, 0]
        elif kind == "zero":
            th[:] = 0
        elif kind == "huge":
            th *= np.float32(1e4)
        elif kind == "tiny":
            th *= np.float32(1e-6)
        want = ref[0].transformer(U, th, (oh, ow))
        assert np.array_equal(C.st_forward(U, th, (oh, ow)), want)
        assert np.array_equal(O.transformer(torch.from_numpy(U), torch.from_numpy(th), (oh, ow)).numpy(), want)
    check()


def test_air_model_source_non_default_sizes():
    """Other constructor sizes (40x40 canvas, 20x20 window, 128 LSTM units, narrower VAE and heads -- the shapes of
    tests/test_gpu_model.py::test_train_parity_non_default_sizes): the source run through the shim and the oracle agree,
    so the oracle's handling of the size hyper-parameters is the reference's."""
    B, cs, ws = 6, 40, 20
    kw = dict(canvas_size=cs, windows_size=ws, rnn_units=128, vae_latent_dimensions=24, vae_recognition_units=(96, 64),
              vae_generative_units=(64, 96), scale_hidden_units=32, shift_hidden_units=32, z_pres_hidden_units=32)
    imgs, cnt = O.synthetic_canvases(B, canvas_size=cs, seed=5, digit_size=(10, 16))
    params = O.init_params(seed=5, **{k: v for k, v in kw.items()})
    params["z_pres/log_odds/output/biases"] += 1.0
    noise = O.make_noise(5, 3, B, latent=24, win=ws * ws)
    for train in (False, True):
        m, _ = _run_source_model(imgs, cnt, params, noise, train=train, **kw)
        out = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=train, **kw).forward(imgs, cnt, noise)
        assert np.array_equal(m.rec_num_digits, out["rec_num_digits"].numpy())
        for k in PER_STEP:
            got = np.asarray(getattr(m, k))
            assert got.shape == tuple(out[k].shape) and _rel(got, out[k].numpy()) < 1e-5, (train, k)
        assert np.asarray(m.reconstruction).shape == (B, cs * cs)


def test_air_model_source_other_hyper_parameters():
    """Every scalar hyper-parameter of the constructor moved off its default (priors, temperature, threshold,
    likelihood std, a fixed instead of an annealed z_pres prior; an annealed schedule with staircase / max): per-step
    outputs, KL terms and the covered-fixture loss of the source and of the oracle agree."""
    from tests import parity_util as PU
    kw = dict(scale_prior_mean=-0.5, scale_prior_variance=0.2, shift_prior_mean=0.1, shift_prior_variance=0.5,
              vae_prior_mean=0.2, vae_prior_variance=2.0, vae_likelihood_std=0.1, z_pres_prior_log_odds=-2.0,
              z_pres_temperature=0.5, stopping_threshold=0.8)
    schedules = (None, {"z_pres_prior_log_odds": {"init": 50.0, "factor": 0.5, "iters": 1000, "staircase": True,
                                                  "max": 20.0, "min": 0.001, "log": True}})
    for annealing in schedules:
        for fixture, train in ((PU.covered_fixture, True), (PU.realistic_fixture, False)):
            imgs, cnt, params, noise = fixture(32, seed=11)
            m, _ = _run_source_model(imgs, cnt, params, noise, train=train, annealing=annealing, global_step=2500, **kw)
            orc = O.AIROracle(params=params, annealing_schedules=annealing, train=train, **kw)
            orc.global_step = 2500
            out = orc.forward(imgs, cnt, noise)
            assert np.array_equal(m.rec_num_digits, out["rec_num_digits"].numpy())
            for k in PER_STEP:
                assert _rel(np.asarray(getattr(m, k)), out[k].numpy()) < 1e-5, (k, train, annealing is None)
            if fixture is PU.covered_fixture:
                assert abs(float(m.loss) - float(out["loss"])) <= 2e-6 * abs(float(out["loss"]))


def test_call_signatures_are_the_references(ref):
    """Drop-in surface: parameter names, order and DEFAULT VALUES of the product's transformer /
    batch_transformer / vae / concrete_* / AIRModel / ModelWrapper equal those of the reference's source, read with
    inspect from the imported reference modules.  The product may only ADD keyword-only / trailing optional arguments."""
    import importlib
    import inspect
    import sys
    import air_b200 as ab
    with S.installed():
        sys.path.insert(0, "/root/reference")
        try:
            ref_model = importlib.import_module("air.air_model").AIRModel
            ref_vae = importlib.import_module("air.vae").vae
            ref_wrapper = importlib.import_module("demo.model_wrapper").ModelWrapper
        finally:
            sys.path.remove("/root/reference")
            for k in [k for k in sys.modules if k.split(".")[0] in ("air", "demo")]:
                del sys.modules[k]

    def params(fn):
        return [(p.name, p.default, p.kind) for p in inspect.signature(fn).parameters.values() if p.name != "self"]

    def same_prefix(ours, theirs, allow_extra=True, skip_defaults=()):
        o, t = params(ours), params(theirs)
        assert len(o) >= len(t), (ours, o, t)
        for (on, od, ok), (tn, td, tk) in zip(o, t):
            assert on == tn and ok == tk, (ours.__name__, on, tn)
            if tn not in skip_defaults:
                assert od == td or (od is inspect.Parameter.empty and td is inspect.Parameter.empty), (ours.__name__, on, od, td)
        for name, default, kind in o[len(t):]:             # additions must be optional
            assert allow_extra and (default is not inspect.Parameter.empty or kind == inspect.Parameter.VAR_KEYWORD), name

    same_prefix(ab.transformer, ref[0].transformer)
    same_prefix(ab.batch_transformer, ref[0].batch_transformer)
    same_prefix(ab.AIRModel.__init__, ref_model.__init__)
    assert [n for n, _, k in params(ab.AIRModel.__init__)[len(params(ref_model.__init__)):]
            if k != inspect.Parameter.KEYWORD_ONLY] == []   # gemm_mode / seed / process_group are keyword-only
    same_prefix(ab.ModelWrapper.__init__, ref_wrapper.__init__, skip_defaults=("session", "data_placeholder"))
    assert params(ab.ModelWrapper.infer)[0][0] == params(ref_wrapper.infer)[0][0] == "images"
    same_prefix(ab.vae, ref_vae, skip_defaults=("activation",))      # default activation: softplus in both, other objects
    for name in ("concrete_binary_sample", "concrete_binary_pre_sigmoid_sample", "concrete_binary_kl_mc_sample"):
        same_prefix(getattr(ab, name), getattr(ref[1], name))

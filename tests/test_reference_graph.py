"""The oracle against the REFERENCE'S OWN GRAPH (no GPU).

tests/golden/ref_graph_*.npz were produced by executing /root/reference/model/air-model.meta -- the MetaGraphDef the
reference's training.py:141 saved: forward loop, TF's autodiff gradient graph, global-norm clip, ApplyAdam, and the
test-mode model -- with the numpy graph interpreter in oracle/tfgraph (make_golden_ref_graph.py).  These tests pin
oracle/air_oracle.py and oracle/st_oracle.c to those outputs; where /root/reference is present (the build container)
the interpreter is also re-run live against the committed files."""
import os

import numpy as np
import pytest
import torch

from oracle import air_oracle as O
from oracle import c_oracle as C
from oracle.tfgraph import air_graph as G
from oracle.tfgraph import pb
from oracle.tfgraph.interp import Interpreter, _strided_index
from tests import parity_util as PU
from tests.golden import make_golden_ref_graph as MG

HAVE_REF = os.path.exists(G.META)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _g(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_st_subgraph_is_bit_exact_with_both_restatements(golden_dir):
    """transformer() as the reference graph wires it (crop 50x50->28x28 and write-back 28x28->50x50, microbench
    poses, ~11 % of the write-back pixels are non-zero out-of-range rounding residues): the C and the torch
    restatements reproduce it BIT FOR BIT, theta and theta^-1 (three separate divisions) included."""
    g = _g(golden_dir, "ref_graph_st.npz")
    s, x, y, imgs, rec = MG.st_inputs()
    B, z = len(s), np.zeros_like(s)
    th = np.stack([s, z, x, z, s, y], 1)
    ti = np.stack([1 / s, z, -x / s, z, 1 / s, -y / s], 1).astype(np.float32)
    assert np.array_equal(g["theta"].reshape(B, 6), th) and np.array_equal(g["theta_inv"].reshape(B, 6), ti)
    assert np.array_equal(C.st_forward(imgs.reshape(B, 50, 50, 1), th, (28, 28))[..., 0], g["crop"])
    assert np.array_equal(C.st_forward(rec.reshape(B, 28, 28, 1), ti, (50, 50))[..., 0], g["back"])
    t = O.transformer(torch.from_numpy(imgs.reshape(B, 50, 50, 1)), torch.from_numpy(th), (28, 28)).numpy()[..., 0]
    assert np.array_equal(t, g["crop"])
    t = O.transformer(torch.from_numpy(rec.reshape(B, 28, 28, 1)), torch.from_numpy(ti), (50, 50)).numpy()[..., 0]
    assert np.array_equal(t, g["back"])
    residues = (np.abs(g["back"]) < 1e-5) & (g["back"] != 0)
    assert residues.mean() > 0.05                         # the fixture does exercise the residue arithmetic


def _check_per_step(out, g, tol):
    for k in ("rec_scales", "rec_shifts", "rec_st_back", "z_pres_probs", "z_pres_kls", "scale_kls", "shift_kls",
              "vae_kls", "running_loss"):
        assert _rel(out[k].numpy().reshape(g[k].shape), g[k]) < tol, (k, _rel(out[k].numpy().reshape(g[k].shape), g[k]))
    assert _rel(out["rec_windows"].numpy()[:, :, ::8], g["rec_windows_sub"]) < tol
    assert np.array_equal(out["rec_num_digits"].numpy(), g["rec_num_digits"])
    assert np.array_equal(out["stop_masks"].numpy(), g["stop_masks"])
    assert out["executed_steps"] == int(g["executed_steps"])
    assert float(out["accuracy"]) == pytest.approx(float(g["accuracy"]), abs=1e-7)


def test_train_step_fp32_covered_fixture(golden_dir):
    """One sess.run(model.training) of the reference graph on the covered fixture (the one the GPU gradient-parity
    test uses): loss, per-step outputs, all 36 raw gradients, the global norm, the clip factor and the Adam-updated
    variables of the oracle agree with the graph's."""
    g = _g(golden_dir, "ref_graph_train_covered.npz")
    imgs, cnt, params, noise = PU.covered_fixture(64, seed=3)
    orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING)
    orc.global_step = 2000
    out, clipped = orc.train_step(imgs, cnt, noise)
    assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    _check_per_step(out, g, 1e-6)
    assert _rel(out["reconstruction_loss"].numpy(), g["reconstruction_loss"]) < 1e-6
    assert _rel(out["reconstruction"].numpy()[:, ::8], g["reconstruction_sub"]) < 1e-6
    assert float(out["grad_global_norm"]) == pytest.approx(float(g["global_norm"]), rel=1e-5)
    scale = float(g["clip_scale"])
    for k, c in clipped.items():                           # train_step returns the clipped gradients
        sample, norm = G.digest(c.numpy(), 1024)
        assert _rel(sample, g["g:" + k] * scale) < 1e-5, k
        assert norm == pytest.approx(float(g["gn:" + k]) * scale, rel=1e-5), k
    for k, v in orc.params.items():                        # ApplyAdam: compare the parameter UPDATE, not the value
        sample, _ = G.digest(v.numpy(), 1024)
        before, _ = G.digest(params[k].numpy(), 1024)
        want = g["v:air/rnn/" + k] - before
        assert _rel(sample - before, want) < 2e-3, (k, _rel(sample - before, want))   # fp32 difference of near-equal numbers
    assert int(g["v:air/global_step"]) == orc.global_step == 2001
    assert float(orc.beta1_power) == float(g["v:air/training/beta1_power"])
    assert float(orc.beta2_power) == float(g["v:air/training/beta2_power"])


def test_train_step_fp64_realistic_poses(golden_dir):
    """Same graph evaluated in fp64 (constants widened to the decimal literals) on realistic poses -- dead steps,
    windows smaller than the canvas, clipping active.  In fp64 the out-of-range residues (~1e-17) are far below the
    1e-9 epsilon of the loss, so structure is compared without the fp32 residue noise."""
    g = _g(golden_dir, "ref_graph_train_realistic_fp64.npz")
    imgs, cnt, params, noise = PU.realistic_fixture(64, seed=1)
    orc = O.AIROracle(params={k: v.double() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING,
                      dtype=torch.float64)
    orc.global_step = 2000
    out, grads = orc.loss_and_grads(imgs.double(), cnt, {k: v.double() for k, v in noise.items()})
    assert abs(float(out["loss"]) - float(g["loss"])) <= 1e-8 * abs(float(g["loss"]))
    _check_per_step(out, g, 1e-6)
    assert len(np.unique(g["rec_num_digits"])) >= 3       # some items stop after 1 or 2 steps
    for k, v in grads.items():
        sample, norm = G.digest(v.numpy(), 1024)
        assert _rel(sample, g["g:" + k]) < 1e-4, (k, _rel(sample, g["g:" + k]))
        assert norm == pytest.approx(float(g["gn:" + k]), rel=1e-4), k


@pytest.mark.parametrize("name,fixture,seed", [("default", PU.default_fixture, 1), ("realistic", PU.realistic_fixture, 2)])
def test_test_mode_graph(golden_dir, name, fixture, seed):
    """The ``air_1`` model (train=False: tf.round on z_pres): digit counts and stopping sums bit-exact, every
    per-step output within 1e-5 (summation order of the GEMMs is the only difference)."""
    g = _g(golden_dir, f"ref_graph_test_{name}.npz")
    imgs, cnt, params, noise = fixture(64, seed=seed)
    orc = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False)
    out = orc.forward(imgs, cnt, noise)
    _check_per_step(out, g, 1e-5)
    assert np.array_equal(out["stopping_sum"].numpy(), g["stopping_sum"])
    # the 88 scalar summaries of air_model.py:160-209, 613-625 (same tags), host logic of demo/model_wrapper.py; the
    # canvas-derived per-item losses are taken from the graph (uncovered fixture)
    import json
    import types
    import air_b200 as ab
    want = json.loads(str(g["summaries"]))
    m = types.SimpleNamespace(**{k: v for k, v in out.items() if torch.is_tensor(v)})
    m.max_digits, m.max_steps = 2, 3
    m.reconstruction_loss, m.loss_per_item = torch.from_numpy(g["reconstruction_loss"]), torch.from_numpy(g["loss_per_item"])
    got = ab.evaluation_summaries(m, cnt)
    assert set(got) == set(want) and len(want) == 88
    for k, v in want.items():
        assert (np.isnan(v) and np.isnan(got[k])) or got[k] == pytest.approx(v, rel=1e-5, abs=1e-6), (k, got[k], v)
    assert len(np.unique(g["rec_num_digits"])) >= 3


def test_st_gradients_bit_exact_with_the_graphs_autodiff(golden_dir):
    """d theta, d theta^-1 and dU as TF's autodiff graph computes them (tapped inside the gradient loop of an fp32
    train step on realistic poses: most canvas pixels out of range, upstream gradients up to 1e7 from the 1e-9
    epsilon of the loss): oracle/st_oracle.c reproduces them BIT FOR BIT."""
    g = _g(golden_dir, "ref_graph_st_grad.npz")
    U = PU.realistic_fixture(64, seed=1)[0][:16].numpy().reshape(16, 50, 50, 1)
    th, ti = g["theta"].reshape(16, 6), g["theta_inv"].reshape(16, 6)
    W = g["window_recon"].reshape(16, 28, 28, 1)
    d_crop, d_wb = g["d_crop"].reshape(16, 28, 28, 1), g["d_writeback"].reshape(16, 50, 50, 1)
    assert np.abs(d_wb).max() > 1e6 and np.abs(th[:, 0] - 1).min() > 0.05
    _, dth = C.st_backward(U, th, d_crop, need_dU=False)
    assert np.array_equal(dth.reshape(16, 2, 3), g["d_theta"])
    dU, dti = C.st_backward(W, ti, d_wb, need_dU=True)
    assert np.array_equal(dti.reshape(16, 2, 3), g["d_theta_inv"])
    assert np.array_equal(dU.reshape(16, 784), g["d_window_recon"])
    # torch restatement (autograd): same formulas, different summation order.  Items whose upstream gradient holds
    # the ~1e7 entries of lit, uncovered pixels are noise-dominated in fp32 in ANY implementation (the graph's own
    # fp32 result is O(1)..O(1e4) away from its fp64 evaluation there), so the tolerance check
    # uses the well-conditioned items; the bit-exact check above covers all of them.
    ok = np.abs(d_wb).reshape(16, -1).max(1) < 10.0
    assert 6 <= ok.sum() < 16
    tW, tti = torch.from_numpy(W).requires_grad_(True), torch.from_numpy(ti).requires_grad_(True)
    O.transformer(tW, tti, (50, 50)).backward(torch.from_numpy(d_wb))
    assert _rel(tti.grad.numpy().reshape(16, 2, 3)[ok], g["d_theta_inv"][ok]) < 1e-5
    assert _rel(tW.grad.numpy().reshape(16, 784)[ok], g["d_window_recon"][ok]) < 5e-4
    tth = torch.from_numpy(th).requires_grad_(True)
    O.transformer(torch.from_numpy(U), tth, (28, 28)).backward(torch.from_numpy(d_crop))
    assert _rel(tth.grad.numpy().reshape(16, 2, 3), g["d_theta"]) < 1e-5


def test_st_gradients_benign_upstream(golden_dir):
    """Same taps, N(0,1) upstream gradient fed into the graph: C restatement bit-exact again; torch autograd shows the
    fp32 conditioning floor of this op (d theta 1e-6, dU ~1e-4: out-of-range pixels add cancelling terms of magnitude
    (x1f - x)(y1f - y) to the edge pixels' dU), which is why the GPU check of dU uses 5e-4."""
    g = _g(golden_dir, "ref_graph_st_grad_benign.npz")
    d_wb, d_crop = (a[:16] for a in MG.benign_upstream())
    U = PU.realistic_fixture(64, seed=1)[0][:16].numpy().reshape(16, 50, 50, 1)
    th, ti, W = g["theta"].reshape(16, 6), g["theta_inv"].reshape(16, 6), g["window_recon"].reshape(16, 28, 28, 1)
    _, dth = C.st_backward(U, th, d_crop.reshape(16, 28, 28, 1), need_dU=False)
    dU, dti = C.st_backward(W, ti, d_wb.reshape(16, 50, 50, 1), need_dU=True)
    assert np.array_equal(dth.reshape(16, 2, 3), g["d_theta"]) and np.array_equal(dti.reshape(16, 2, 3), g["d_theta_inv"])
    assert np.array_equal(dU.reshape(16, 784), g["d_window_recon"])
    tW, tti = torch.from_numpy(W).requires_grad_(True), torch.from_numpy(ti).requires_grad_(True)
    O.transformer(tW, tti, (50, 50)).backward(torch.from_numpy(d_wb.reshape(16, 50, 50, 1)))
    assert _rel(tti.grad.numpy().reshape(16, 2, 3), g["d_theta_inv"]) < 1e-5
    assert _rel(tW.grad.numpy().reshape(16, 784), g["d_window_recon"]) < 5e-4


def test_five_step_inference_graph(golden_dir):
    """BASELINE configs[4] runs 5 attention steps.  The reference graph uses max_steps in exactly one place inside the
    loop, the constant of cond()'s ``step < max_steps``; fed as 5, the same serialized graph is the 5-step model."""
    g = _g(golden_dir, "ref_graph_test_realistic_T5.npz")
    imgs, cnt, params, noise = PU.realistic_fixture(64, seed=8, T=5)
    out = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False, max_steps=5).forward(imgs, cnt, noise)
    assert g["rec_scales"].shape == (64, 5, 1) and int(g["executed_steps"]) == 5
    _check_per_step(out, g, 1e-5)
    assert np.array_equal(out["stopping_sum"].numpy(), g["stopping_sum"])
    assert set(np.unique(g["rec_num_digits"])) == {0, 1, 2, 3, 4, 5}


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_interpreter_reproduces_committed_goldens_live(golden_dir):
    nodes = pb.load_metagraph(G.META)
    assert nodes[1] == "1.3.0" and len(nodes[0]) == 15022
    s, x, y, imgs, rec = MG.st_inputs()
    win, back, _, _ = MG.graph_st(nodes, s, x, y, imgs, rec)
    g = _g(golden_dir, "ref_graph_st.npz")
    assert np.array_equal(win, g["crop"]) and np.array_equal(back, g["back"])
    imgs, cnt, params, noise = PU.covered_fixture(64, seed=3)
    out = G.run_train_step(nodes, params, imgs, cnt, noise, global_step=2000)
    g = _g(golden_dir, "ref_graph_train_covered.npz")
    assert float(out["loss"]) == float(g["loss"]) and out["executed_steps"] == 3
    assert np.array_equal(G.digest(out["raw_grads"]["rnn/kernel"], 1024)[0], g["g:rnn/kernel"])
    assert out["noise_used"] == {(k, t) for k in ("scale", "shift", "vae_latent", "vae_like", "concrete_u")
                                 for t in range(3)}


def test_interpreter_pieces():
    """Unit checks of the interpreter's own machinery on hand-built graphs: a while loop with a stack-based
    'gradient' loop, StridedSlice masks, BroadcastGradientArgs."""
    class N:
        def __init__(self, name, op, inputs=(), ctrl=(), **attr):
            self.name, self.op, self.ctrl, self.attr = name, op, list(ctrl), attr
            self.inputs = [(i.partition(":")[0], int(i.partition(":")[2] or 0)) for i in inputs]
    c = lambda name, v: N(name, "Const", value=np.asarray(v))                       # noqa: E731
    fr = dict(frame_name=b"f")
    nodes = [c("zero", np.int32(0)), c("one", np.int32(1)), c("n", np.int32(4)), c("x0", np.float32(1.5)),
             N("stk", "Stack"),
             N("e_i", "Enter", ["zero"], **fr), N("e_x", "Enter", ["x0"], **fr), N("e_n", "Enter", ["n"], **fr),
             N("e_one", "Enter", ["one"], **fr), N("e_s", "RefEnter", ["stk"], **fr),
             N("m_i", "Merge", ["e_i", "nx_i"]), N("m_x", "Merge", ["e_x", "nx_x"]),
             N("lt", "Less", ["m_i", "e_n"]), N("lc", "LoopCond", ["lt"]),
             N("s_i", "Switch", ["m_i", "lc"]), N("s_x", "Switch", ["m_x", "lc"]),
             N("id_i", "Identity", ["s_i:1"]), N("id_x", "Identity", ["s_x:1"]),
             N("push", "StackPush", ["e_s", "id_x"]),
             N("add", "Add", ["id_i", "e_one"], ctrl=["push"]), N("sq", "Mul", ["id_x", "id_x"]),
             N("nx_i", "NextIteration", ["add"]), N("nx_x", "NextIteration", ["sq"]),
             N("x_i", "Exit", ["s_i"]), N("x_x", "Exit", ["s_x"])]
    # second loop: pops the stashed x values in reverse and multiplies 2*x into an accumulator (d/dx0 of x^(2^n))
    fb = dict(frame_name=b"b")
    nodes += [c("onef", np.float32(1.0)), c("two", np.float32(2.0)),
              N("b_c", "Enter", ["x_i"], **fb), N("b_g", "Enter", ["onef"], **fb), N("b_s", "RefEnter", ["stk"], **fb),
              N("b_two", "Enter", ["two"], **fb), N("b_one", "Enter", ["one"], **fb),
              N("bm_c", "Merge", ["b_c", "bn_c"]), N("bm_g", "Merge", ["b_g", "bn_g"]),
              N("ge", "GreaterEqual", ["bm_c", "b_one"]), N("blc", "LoopCond", ["ge"]),
              N("bs_c", "Switch", ["bm_c", "blc"]), N("bs_g", "Switch", ["bm_g", "blc"]),
              N("bi_c", "Identity", ["bs_c:1"]), N("bi_g", "Identity", ["bs_g:1"]),
              N("pop", "StackPop", ["b_s"], ctrl=["bi_c"]), N("tx", "Mul", ["b_two", "pop"]),
              N("g2", "Mul", ["bi_g", "tx"]), N("dec", "Sub", ["bi_c", "b_one"]),
              N("bn_c", "NextIteration", ["dec"]), N("bn_g", "NextIteration", ["g2"]),
              N("bx_g", "Exit", ["bs_g"])]
    I = Interpreter(({n.name: n for n in nodes}, "unit"))
    assert int(I.fetch("x_i")) == 4 and float(I.fetch("x_x")) == pytest.approx(1.5 ** 16)
    assert float(I.fetch("bx_g")) == pytest.approx(16 * 1.5 ** 15)      # d/dx x^16
    assert I.trip_count("f") == 4 and I.trip_count("b") == 4

    a = np.arange(24).reshape(2, 3, 4)
    at = dict(begin_mask=0, end_mask=0, ellipsis_mask=0, new_axis_mask=0, shrink_axis_mask=0)
    assert np.array_equal(a[_strided_index(a.shape, [0, 1, 0], [2, 3, 4], [1, 1, 2], at)], a[0:2, 1:3, 0:4:2])
    assert np.array_equal(a[_strided_index(a.shape, [0, 2], [0, 3], [1, 1], dict(at, begin_mask=1, end_mask=1, shrink_axis_mask=2))],
                          a[:, 2])
    assert np.array_equal(a[_strided_index(a.shape, [0, 1], [0, 2], [1, 1], dict(at, ellipsis_mask=1, shrink_axis_mask=2))],
                          a[..., 1])
    r0, r1 = I.op_BroadcastGradientArgs(None, None, np.array([64, 1]), np.array([64, 784]))
    assert list(r0) == [1] and list(r1) == []
    r0, r1 = I.op_BroadcastGradientArgs(None, None, np.array([], np.int32), np.array([64]))
    assert list(r0) == [0] and list(r1) == []


# z_pres_prior_log_odds as the reference graph computes it from global_step (exponential_decay -> maximum -> +1e-9 ->
# log, air_model.py:94-121 with training.py:110-115's schedule), float32 values in hex
ANNEALED = {0: '0x1.26bb1c0000000p+3', 1: '0x1.26b4d20000000p+3', 777: '0x1.13a5a60000000p+3',
            2999: '0x1.ba253c0000000p+2', 3000: '0x1.ba18aa0000000p+2', 12000: '0x0.0p+0',
            30000: '-0x1.ba107a0000000p+3', 40000: '-0x1.407b5e0000000p+4', 270000: '-0x1.407b5e0000000p+4'}


def test_annealing_schedule_bit_exact_with_graph():
    nodes = pb.load_metagraph(G.META) if HAVE_REF else None
    for gs, hx in ANNEALED.items():
        want = np.float32(float.fromhex(hx))
        got = O.annealed_value(O.DEFAULT_ANNEALING["z_pres_prior_log_odds"], gs)
        assert np.float32(got.item()) == want, gs
        if nodes is not None:
            live = Interpreter(nodes, {"air/global_step": np.int32(gs)}).fetch("air/z_pres_prior_log_odds_log")
            assert np.float32(live) == want, gs


def test_reconstruction_image_summary_bit_exact(golden_dir):
    """demo/visualize.py (resize x2, window frames through the write-back ST with rec_st_back, R/G/B per step) equals
    the tensor the reference graph feeds to tf.summary.image("reconstruction") bit for bit; the ST is the oracle's
    here and the CUDA kernel's in tests/test_gpu_zz_reference_graph.py."""
    import hashlib
    import air_b200 as ab
    g = _g(golden_dir, "ref_graph_vis.npz")
    imgs = PU.realistic_fixture(64, seed=2)[0][:12]
    got = ab.visualize_reconstructions(imgs, torch.from_numpy(g["reconstruction"]), torch.from_numpy(g["rec_st_back"]),
                                       torch.from_numpy(g["rec_num_digits"]), transformer=O.transformer).numpy()
    assert got.shape == (12, 100, 204, 3) and got.min() == 0.0 and got.max() == 1.0
    assert np.array_equal(got[:4], g["image_full"])
    assert [hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() for a in got] == list(g["sha256"])
    assert len(set(g["rec_num_digits"].tolist())) >= 3       # frames of 0, 1, 2 and 3 executed steps are all drawn
    # fewer executed steps than max_steps: the missing matrices are zero-padded (air_model.py:227-231)
    short = ab.visualize_reconstructions(imgs, torch.from_numpy(g["reconstruction"]),
                                         torch.from_numpy(g["rec_st_back"][:, :1]),
                                         torch.clamp(torch.from_numpy(g["rec_num_digits"]), max=1),
                                         transformer=O.transformer)
    assert short.shape == (12, 100, 204, 3) and not torch.equal(short, torch.from_numpy(got))


def test_gpu_reference_graph_tests_rehearsed_on_the_oracle(golden_dir, monkeypatch):
    """The GPU tests of tests/test_gpu_zz_reference_graph.py, run here with the CUDA model replaced by an
    oracle-backed stand-in: proves that their fixtures, global steps, attribute names and tolerances are consistent
    with the committed vectors before they ever reach a GPU box."""
    import types
    import tests.test_gpu_zz_reference_graph as T

    class Stand:
        def __init__(self, orc, imgs, cnt):
            self.orc, self.imgs, self.cnt = orc, imgs, cnt
            self.store = types.SimpleNamespace(named_grads=lambda: self.grads)

        def _take(self, out):
            for k, v in out.items():
                setattr(self, k, v)

        def loss_and_grads(self, noise):
            out, self.grads = self.orc.loss_and_grads(self.imgs, self.cnt, noise)
            self._take(out)

        def run(self, noise):
            self._take(self.orc.forward(self.imgs, self.cnt, noise))

    def make_pair(imgs, cnt, params, train=True, global_step=2000, gemm_mode="fp32", **kw):
        orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING,
                          train=train, **kw)
        orc.global_step = global_step
        return orc, Stand(orc, imgs, cnt)

    monkeypatch.setattr(T, "make_pair", make_pair)
    monkeypatch.setattr(T, "cuda_noise", lambda n: n)
    monkeypatch.setattr(T, "_cu", lambda a: torch.from_numpy(np.ascontiguousarray(a)))
    monkeypatch.setattr(T.ab, "transformer", O.transformer)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    T.test_st_kernels_bit_exact_with_reference_graph(golden_dir)
    T.test_train_step_against_reference_graph(golden_dir, "fp32")
    T.test_test_mode_against_reference_graph(golden_dir, "default", PU.default_fixture, 1, "fp32")
    T.test_test_mode_against_reference_graph(golden_dir, "realistic", PU.realistic_fixture, 2, "fp32")
    T.test_train_mode_realistic_poses_against_fp64_graph_run(golden_dir, "fp32")
    T.test_five_step_inference_against_reference_graph(golden_dir, "fp32")
    T.test_st_backward_kernels_against_the_graphs_autodiff(golden_dir)
    from tests.test_trained_weights import PATH as trained_path
    if os.path.exists(trained_path):
        T.test_trained_weights_inference_counts_digits()
    orig = T.ab.visualize_reconstructions
    monkeypatch.setattr(T.ab, "visualize_reconstructions", lambda *a, **k: orig(*a, transformer=O.transformer, **k))
    T.test_reconstruction_image_summary_bit_exact_on_device(golden_dir)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_three_consecutive_train_steps_follow_the_graph():
    """sess.run(model.training) three times, variables / Adam slots / beta powers / global_step carried from one run
    of the graph to the next, fresh noise per step: the oracle's losses follow to 1e-6 and its accumulated parameter
    movement to 2e-4 of that movement (annealed prior, clip, ApplyAdam and the beta-power updates all in the loop)."""
    nodes = pb.load_metagraph(G.META)
    imgs, cnt, params, _ = PU.covered_fixture(64, seed=5)
    orc = O.AIROracle(params={k: v.clone() for k, v in params.items()}, annealing_schedules=O.DEFAULT_ANNEALING)
    P = {k: v.numpy().copy() for k, v in params.items()}
    state = dict(adam_m=None, adam_v=None, global_step=0, beta1_power=np.float32(0.9), beta2_power=np.float32(0.999))
    for step in range(3):
        nz = O.make_noise(100 + step, 3, 64)
        nv = (out := G.run_train_step(nodes, P, imgs, cnt, nz, **state))["new_variables"]
        P = {k: nv["air/rnn/" + k] for k in P}
        state = dict(adam_m={k: nv[f"air/training/air/rnn/{k}/Adam"] for k in P},
                     adam_v={k: nv[f"air/training/air/rnn/{k}/Adam_1"] for k in P},
                     global_step=int(nv["air/global_step"]), beta1_power=nv["air/training/beta1_power"],
                     beta2_power=nv["air/training/beta2_power"])
        o, _ = orc.train_step(imgs, cnt, nz)
        assert abs(float(o["loss"]) - float(out["loss"])) <= 1e-6 * abs(float(out["loss"]))
        assert state["global_step"] == orc.global_step == step + 1
        assert float(state["beta1_power"]) == float(orc.beta1_power)
        assert float(state["beta2_power"]) == float(orc.beta2_power)
        for k in P:
            moved = np.linalg.norm(P[k] - params[k].numpy())
            if moved > 0:
                assert np.linalg.norm(P[k] - orc.params[k].numpy()) <= 2e-4 * moved, (step, k)


def test_protobuf_decoder_on_hand_assembled_messages():
    """oracle/tfgraph/pb.py against bytes assembled by hand from the .proto field numbers (NodeDef with data and
    control inputs, tensor / int / shape / list attributes, negative varints, the repeat-last-value rule)."""
    import struct

    def vint(v):
        v &= (1 << 64) - 1
        out = b""
        while True:
            b = v & 0x7F
            v >>= 7
            out += bytes([b | (0x80 if v else 0)])
            if not v:
                return out

    def ld(fn, payload):
        return vint(fn << 3 | 2) + vint(len(payload)) + payload

    def vf(fn, v):
        return vint(fn << 3) + vint(v)
    dim = lambda n: ld(2, vf(1, n))                                                  # noqa: E731
    tensor = vf(1, 1) + ld(2, dim(2) + dim(3)) + ld(5, struct.pack("<2f", 1.5, -2.0))   # float [2,3]: 1.5, -2, -2 ...
    itensor = vf(1, 3) + ld(2, b"") + ld(7, vint(-7))                                # int32 scalar -7
    content = vf(1, 9) + ld(2, dim(2)) + ld(4, struct.pack("<2q", 5, -6))            # int64 [2] in tensor_content
    attr = lambda k, v: ld(5, ld(1, k.encode()) + ld(2, v))                          # noqa: E731
    node = (ld(1, b"scope/n") + ld(2, b"Const") + ld(3, b"a:1") + ld(3, b"b") + ld(3, b"^c") +
            attr("value", ld(8, tensor)) + attr("ival", ld(8, itensor)) + attr("cval", ld(8, content)) +
            attr("i", vf(3, -3)) + attr("f", vint(4 << 3 | 5) + struct.pack("<f", 0.25)) + attr("b", vf(5, 1)) +
            attr("s", ld(2, b"frame")) + attr("shape", ld(7, dim(-1) + dim(2500))) +
            attr("list", ld(1, ld(3, vint(1) + vint(2) + vint(-1)))))
    nd = pb.Node(memoryview(node))
    assert (nd.name, nd.op, nd.inputs, nd.ctrl) == ("scope/n", "Const", [("a", 1), ("b", 0)], ["c"])
    assert np.array_equal(nd.attr["value"], np.array([[1.5, -2, -2], [-2, -2, -2]], np.float32))
    assert nd.attr["ival"].dtype == np.int32 and int(nd.attr["ival"]) == -7 and nd.attr["ival"].shape == ()
    assert nd.attr["cval"].dtype == np.int64 and list(nd.attr["cval"]) == [5, -6]
    assert nd.attr["i"] == -3 and nd.attr["f"] == 0.25 and nd.attr["b"] is True and nd.attr["s"] == b"frame"
    assert nd.attr["shape"] == ("shape", [-1, 2500]) and nd.attr["list"] == [1, 2, -1]


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_fused_canvas_update_bit_exact_with_the_graph(golden_dir):
    """ST write-back + z_pres scaling + stop-mask select + accumulation (air_model.py:363-366, 429-439) as the graph
    wires them, fed with the committed st.npz operands (incl. stopping sums exactly AT the threshold): the C
    restatement of the fused op (what air_st_writeback_canvas_fwd is tested against) is bit-exact."""
    g = _g(golden_dir, "st.npz")
    n = len(g["z"])
    assert n <= 64
    pad = lambda a: np.concatenate([a, np.zeros((64 - n,) + a.shape[1:], a.dtype)])     # noqa: E731  (graph batch = 64)
    window_st = pad(g["back"].reshape(n, 50, 50))         # ST(window, theta_inv), itself bit-exact (test above)
    nodes = pb.load_metagraph(G.META)
    I = Interpreter(nodes, {}, {"air/rnn/while/st_backward/strided_slice:0": window_st,
                                "air/rnn/while/z_pres/gumbel/Sigmoid:0": pad(g["z"]),
                                "air/rnn/while/add:0": pad(g["stop"]),
                                "air/rnn/while/Identity_4:0": pad(g["canvas"])})
    got = I.eval("air/rnn/while/canvas/add", 0, 0)[:n]
    assert np.array_equal(got, g["canvas_out"])
    assert np.array_equal(got, C.canvas_update(g["canvas"], g["back"].reshape(-1, 2500), g["z"], g["stop"], 0.99))
    assert (g["stop"] == np.float32(0.99)).any()          # 0.99 itself is NOT < 0.99: those rows stay untouched


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("scope,train", [("air", 1), ("air_1", 0)])
def test_concrete_act_step_subgraph(golden_dir, scope, train):
    """Concrete sample, its MC KL against the (annealed) prior, the z_pres KL masked by the OLD stopping sum, the new
    stopping sum and the digit counter, evaluated inside the reference graph's loop body for fed log-odds / uniform
    noise / loop state (train graph: continuous z; test graph: tf.round): the C restatement of the fused step (what
    air_concrete_step_fwd is tested against) gives the same integers and masks exactly and the floats to an ulp of
    log/exp."""
    g = _g(golden_dir, "concrete.npz")
    w = f"{scope}/rnn/while/"
    nodes = pb.load_metagraph(G.META)
    I = Interpreter(nodes, {}, {w + "z_pres/log_odds/output/strided_slice:0": g["log_odds"],
                                w + "Identity_1:0": g["stop_prev"], w + "Identity_5:0": g["loss_prev"],
                                w + "Identity_6:0": g["digits_prev"],
                                f"{scope}/z_pres_prior_log_odds_log:0": np.float32(g["prior"])},
                    random_fn=lambda node, it, shape: g["u"])
    want = C.concrete_step(g["log_odds"], g["u"], g["stop_prev"], g["loss_prev"], g["digits_prev"], float(g["prior"]),
                           1.0, 0.99, train)
    z_node = "z_pres/gumbel/Round" if not train else "z_pres/gumbel/Sigmoid"
    got = {"y": I.eval(w + "z_pres/gumbel/truediv", 0, 0), "z": I.eval(w + z_node, 0, 0),
           "kl": I.eval(w + "loss/z_pres_kl/sub_4", 0, 0), "loss_new": I.eval(w + "loss/z_pres_kl/add_8", 0, 0),
           "stop_new": I.eval(w + "add", 0, 0), "digits_new": I.eval(w + "add_1", 0, 0)}
    assert np.array_equal(got["digits_new"], want["digits_new"])
    assert np.array_equal(got["stop_new"] < np.float32(0.99), want["stop_new"] < np.float32(0.99))
    assert np.array_equal(got["loss_new"] != g["loss_prev"], want["loss_new"] != g["loss_prev"])     # same KL mask
    if not train:
        assert np.array_equal(got["z"], want["z"]) and np.array_equal(got["stop_new"], want["stop_new"])
    for k in ("y", "z", "kl", "loss_new", "stop_new"):
        np.testing.assert_allclose(got[k], want[k], rtol=2e-6, atol=2e-6, err_msg=k)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present (GPU box)")
def test_reconstruction_loss_and_its_gradient_tie_rules():
    """air_model.py:580-593 and the gradient TF generated for it, with canvas values exactly at and beyond the clip
    bounds: the loss vector, and d(mean loss)/d(canvas) -- MinimumGrad passes where canvas <= 1, MaximumGrad where the
    clipped value >= 0, so the gradient passes AT 0.0 and AT 1.0 and is blocked outside -- equal the oracle's."""
    rng = np.random.default_rng(9)
    canvas = rng.uniform(-0.2, 1.2, (64, 2500)).astype(np.float32)
    canvas.ravel()[::7] = 0.0
    canvas.ravel()[3::11] = 1.0
    canvas.ravel()[5::13] = np.float32(-1e-7)             # a negative rounding residue
    x = (rng.uniform(0, 1, (64, 2500)) * (rng.uniform(0, 1, (64, 2500)) < 0.3)).astype(np.float32)
    nodes = pb.load_metagraph(G.META)
    I = Interpreter(nodes, {}, {"air/rnn/while/Exit_4:0": canvas, "pipeline/shuffle_batch:0": x,
                                "air/rnn/while/Exit_5:0": np.zeros(64, np.float32)})
    rec_loss = I.fetch("air/loss/reconstruction/Neg")
    dcanvas = I.fetch("air/training/gradients/air/loss/reconstruction/Minimum_grad/tuple/control_dependency")
    tc = torch.from_numpy(canvas).requires_grad_(True)
    tx = torch.from_numpy(x)
    r = torch.clamp(tc, 0.0, 1.0)                         # the oracle's statement of max(min(r, 1), 0)  (:582)
    loss_vec = -torch.sum(tx * torch.log(r + O.EPS) + (1.0 - tx) * torch.log(1.0 - r + O.EPS), 1)
    torch.mean(loss_vec).backward()
    np.testing.assert_allclose(loss_vec.detach().numpy(), rec_loss, rtol=2e-6)
    got, want = dcanvas, tc.grad.numpy()
    assert np.array_equal(got != 0, want != 0)            # same pass / block pattern
    assert (got[canvas == 0.0] != 0).all() and (got[canvas == 1.0] != 0).any()
    assert (got[canvas < 0] == 0).all() and (got[canvas > 1] == 0).all()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-8)      # atol: the two terms of the derivative cancel

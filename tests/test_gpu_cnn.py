"""GPU parity tests of the CNN front-end (AIRModel(cnn=True), air_model.py:510-535): the conv kernels through the
C ABI against torch-CPU autograd of the same ops, and the full model against the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import air_b200 as ab
from air_b200 import checkpoint, data, ops
from oracle import air_oracle as O
from tests.parity_util import covered_fixture, cuda_noise, make_pair, relnorm

pytestmark = pytest.mark.gpu
DEV = "cuda"
LAYERS = [(1, 50, 50, True), (8, 25, 25, True), (8, 12, 12, False)]   # (cin, H, W, pool) of conv1..3


def _torch_layer(x_nhwc, w_hwio, b, pool):
    t = torch.relu(F.conv2d(x_nhwc.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), b, padding=2))
    if pool:
        t = F.max_pool2d(t, 2, 2)
    return t.permute(0, 2, 3, 1)


@pytest.mark.parametrize("F_", [8, 4, 16])
@pytest.mark.parametrize("cin,H,W,pool", LAYERS)
@pytest.mark.parametrize("B", [1, 37])
def test_conv_layer_forward_backward(cin, H, W, pool, B, F_):
    """F_ = cnn_filters (air_model.py:21, default 8): output channels of every layer, input channels of conv2 / conv3."""
    if cin > 1:
        cin = F_
    g = torch.Generator().manual_seed(100 + cin + H + B)
    x = torch.rand(B, H, W, cin, generator=g) * (torch.rand(B, H, W, cin, generator=g) > 0.6)   # sparse, like a canvas
    w = (torch.rand(5, 5, cin, F_, generator=g) - 0.45) * 0.5
    b = (torch.rand(F_, generator=g) - 0.5) * 0.2
    xt, wt, bt = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    want = _torch_layer(xt, wt, bt, pool)
    PH, PW = (H // 2, W // 2) if pool else (H, W)
    out = torch.empty(B, PH, PW, F_, device=DEV)
    arg = torch.empty(B, PH, PW, F_, device=DEV, dtype=torch.uint8) if pool else None
    xg, wg, bg = x.to(DEV), w.to(DEV), b.to(DEV)
    ops.conv5x5_fwd(xg, wg, bg, out, arg, H, W, cin, F_, pool)
    np.testing.assert_allclose(out.cpu().numpy(), want.detach().numpy(), rtol=2e-5, atol=2e-6)
    assert (want > 0).float().mean() > 0.2                      # the test exercises live units
    dout = torch.randn(B, PH, PW, F_, generator=g)
    want.backward(dout)
    dw, db = torch.full((5, 5, cin, F_), 7.0, device=DEV), torch.full((F_,), 7.0, device=DEV)
    din = torch.empty(B, H, W, cin, device=DEV) if cin > 1 else None
    ws = torch.zeros(ops.conv5x5_bwd_workspace(B, cin, F_), device=DEV)
    ops.conv5x5_bwd(xg, wg, out, arg, dout.to(DEV), din, dw, db, False, ws, H, W, cin, F_, pool)
    assert relnorm(dw, wt.grad) < 1e-5 and relnorm(db, bt.grad) < 1e-5
    if din is not None:
        assert relnorm(din, xt.grad) < 1e-5
    # accumulate == 1 adds to the existing gradient; the run is bit-reproducible
    dw2, db2 = dw.clone(), db.clone()
    ops.conv5x5_bwd(xg, wg, out, arg, dout.to(DEV), din, dw2, db2, True, ws, H, W, cin, F_, pool)
    assert torch.equal(dw2, dw + dw) and torch.equal(db2, db + db)


def test_conv_unsupported_shape_fails_loudly():
    x = torch.zeros(2, 10, 10, 3, device=DEV)
    with pytest.raises(ab.AirError, match="only the three layers"):
        ops.conv5x5_fwd(x, torch.zeros(5, 5, 3, 8, device=DEV), torch.zeros(8, device=DEV), torch.zeros(2, 10, 10, 8, device=DEV),
                        None, 10, 10, 3, 8, False)


@pytest.mark.parametrize("filters", [8, 4, 16])
def test_model_cnn_forward_and_gradients_vs_oracle(filters):
    imgs, cnt, params, noise = covered_fixture(8, seed=2, cnn=True)
    if filters != 8:   # (the fixture's parameter dict is built for the default 8 filters: re-draw the front-end and the LSTM input rows)
        fresh = O.init_params(seed=2, cnn=True, cnn_filters=filters)
        for k in fresh:
            if k.startswith("cnn/") or k == "rnn/kernel":
                params[k] = fresh[k]
    orc, m = make_pair(imgs, cnt, params, train=True, cnn=True, cnn_filters=filters)
    assert m.Kx.shape == (144 * filters, 1024)
    out, grads = orc.loss_and_grads(imgs, cnt, noise)
    m.loss_and_grads(cuda_noise(noise))
    assert torch.equal(m.rec_num_digits.cpu(), out["rec_num_digits"])
    assert abs(m.loss.item() - out["loss"].item()) / abs(out["loss"].item()) < 1e-5
    got = m.store.named_grads()
    assert set(got) == set(grads) and any(k.startswith("cnn/") for k in got)
    for k, gcu in got.items():
        assert relnorm(gcu, grads[k]) < 1e-4, (k, relnorm(gcu, grads[k]))
    for k in ("cnn/conv1/kernel", "cnn/conv2/kernel", "cnn/conv3/kernel", "cnn/conv3/bias"):
        assert float(grads[k].abs().max()) > 0                 # the gradient really reaches the front-end


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_model_cnn_trains_graph_replay_and_checkpoint_names(mode, tmp_path):
    imgs, cnt = O.synthetic_canvases(64, seed=5)
    ab.reset_variable_scopes()
    hyper = dict(data.TRAINING_HYPER, cnn=True)
    m = ab.AIRModel(imgs.cuda(), cnt.cuda(), train=True, gemm_mode=mode, seed=1, **hyper)
    m.capture()
    losses = []
    for _ in range(30):
        m.train_step()
        losses.append(m.loss.item())
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])
    names = checkpoint.model_tensors(m.store)
    assert "air/cnn/conv1/kernel" in names and names["air/cnn/conv2/kernel"].shape == (5, 5, 8, 8)
    assert "air/training/air/cnn/conv3/bias/Adam_1" in names and names["air/rnn/rnn/kernel"].shape == (1408, 1024)
    m.save(str(tmp_path / "ck"))
    ab.reset_variable_scopes()
    m2 = ab.AIRModel(imgs.cuda(), cnt.cuda(), train=False, gemm_mode=mode, seed=9, **hyper)
    m2.restore(str(tmp_path / "ck"))
    for k, v in m.store.named_views().items():
        assert torch.equal(v, m2.store.named_views()[k]), k
    m2.run()
    assert m2.rec_num_digits.shape == (64,)

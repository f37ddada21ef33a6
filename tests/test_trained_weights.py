"""Trained weights fixture (no GPU).  The reference's checkpoint blob (model/air-model.data-00000-of-00001) is not in
the checkout, so tests/golden/trained_synthetic_fp16.npz holds the stand-in: the parameters after the 25,000-iteration
oracle run of oracle/train_convergence.py (reference training configuration, synthetic canvases), rounded to float16
to halve the file.  Everything that consumes them widens the SAME float16 values to float32, so CPU and GPU see
identical weights."""
import os

import numpy as np
import pytest
import torch

from oracle import air_oracle as O
from tests.conftest import GOLDEN

PATH = os.path.join(GOLDEN, "trained_synthetic_fp16.npz")
needs_weights = pytest.mark.skipif(not os.path.exists(PATH), reason="trained-weights fixture not generated")


def trained_params(path=PATH):
    g = np.load(path)
    return {k: torch.from_numpy(g[k].astype(np.float32)) for k in g.files if k != "global_step"}, int(g["global_step"])


def heldout(B=256, seed=777, T=3):
    imgs, cnt = O.synthetic_canvases(B, seed=seed)
    return imgs, cnt, O.make_noise(seed, T, B)


@needs_weights
def test_trained_oracle_counts_digits():
    params, step = trained_params()
    assert step == 25000 and set(params) == set(O.init_params(seed=0))
    imgs, cnt, noise = heldout()
    orc = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False)
    orc.global_step = step
    out = orc.forward(imgs, cnt, noise)
    hit = (out["rec_num_digits"] == cnt).float()
    assert hit.mean().item() >= 0.93, hit.mean().item()
    for k in range(3):
        assert hit[cnt == k].mean().item() >= 0.85, (k, hit[cnt == k].mean().item())
    # the attention windows sit on the blobs: the live steps' reconstruction explains the image far better than an
    # empty canvas would
    empty = -(imgs * np.log(1e-9) + (1 - imgs) * np.log(1 + 1e-9)).sum(1)
    assert out["reconstruction_loss"][cnt > 0].mean() < 0.25 * empty[cnt > 0].mean()


@needs_weights
def test_trained_weights_survive_the_tf_bundle_format(tmp_path):
    """save_checkpoint / load_checkpoint (the tf.train.Saver bundle format of training.py:141) with real trained
    values and the reference's variable names."""
    from air_b200 import checkpoint
    params, step = trained_params()
    tensors = {"air/rnn/" + k: v.numpy() for k, v in params.items()}
    tensors["air/global_step"] = np.int32(step)
    checkpoint.save_checkpoint(str(tmp_path / "air-model-25000"), tensors)
    back = checkpoint.load_checkpoint(str(tmp_path / "air-model-25000"))
    assert set(back) == set(tensors) and int(back["air/global_step"]) == 25000
    for k, v in tensors.items():
        assert np.array_equal(back[k], v), k


@needs_weights
@pytest.mark.skipif(not os.path.exists("/root/reference/model/air-model.meta"), reason="/root/reference not present")
def test_reference_graph_counts_digits_with_the_trained_weights():
    """The reference's own test-mode graph, loaded with the trained weights and run by the interpreter, infers the
    same digit counts as the oracle on held-out canvases and is right about them (its own accuracy summary)."""
    from oracle.tfgraph import air_graph as G
    from oracle.tfgraph import pb
    params, step = trained_params()
    imgs, cnt, noise = heldout(B=64, seed=778)
    out = G.run_test_model(pb.load_metagraph(G.META), params, imgs, cnt, noise)
    orc = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False)
    o = orc.forward(imgs, cnt, noise)
    assert np.array_equal(out["rec_num_digits"], o["rec_num_digits"].numpy())
    assert np.array_equal(out["stop_masks"], o["stop_masks"].numpy())
    assert float(out["accuracy"]) >= 0.9 and out["summaries"]["digit_acc_all_dig"] == pytest.approx(float(out["accuracy"]))
    for k in ("rec_scales", "rec_shifts", "rec_windows"):
        a, b = out[k], o[k].numpy().reshape(out[k].shape)
        assert np.linalg.norm(a - b) <= 1e-5 * np.linalg.norm(b), k

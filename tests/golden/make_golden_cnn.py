"""Golden vector of the CNN front-end restatement (oracle.cnn_frontend, air_model.py:510-535): inputs, the six
conv tensors, the [B,1152] feature map and the parameter gradients of a fixed linear functional of it.
The cnn=True front-end is not in the reference's saved graph: its wiring is pinned by executing the source through the
tf shim (tests/test_reference_source.py); this file is an oracle regression pin of the convolution arithmetic (torch
conv2d, cross-checked against plain-C loops), not a reference output.

    python tests/golden/make_golden_cnn.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import air_oracle as O  # noqa: E402


def build():
    imgs, _ = O.synthetic_canvases(3, seed=11)
    p = {k: v.clone() for k, v in O.init_params(seed=7, cnn=True).items() if k.startswith("cnn/")}
    g = torch.Generator().manual_seed(5)
    for k in p:
        if k.endswith("bias"):
            p[k] = (torch.rand(p[k].shape, generator=g) - 0.4) * 0.1
    G = torch.randn(3, 1152, generator=g)
    leaf = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    feat = O.cnn_frontend(imgs, leaf)
    (feat * G).sum().backward()
    out = {"x": imgs.numpy(), "G": G.numpy(), "features": feat.detach().numpy()}
    for k, v in p.items():
        out["p:" + k] = v.numpy()
        out["g:" + k] = leaf[k].grad.numpy()
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "cnn.npz"), **build())
    print("wrote cnn.npz")

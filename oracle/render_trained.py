"""Qualitative check (test infrastructure, CPU): attention windows and reconstructions of the TRAINED oracle on held-out
synthetic canvases, drawn with the product's demo/visualize.py (the reference's image summary, air_model.py:211-267)
driven by the oracle's Spatial Transformer -- the counterpart of the reference's images/rec_samples.png.

    python oracle/render_trained.py tests/golden/trained_synthetic_fp16.npz profiles/r1_trained_reconstructions.png
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import air_oracle as O                      # noqa: E402
from tests.test_trained_weights import heldout, trained_params   # noqa: E402


def main(weights, out_png, n=24, cols=4):
    from PIL import Image
    import air_b200 as ab
    params, step = trained_params(weights)
    imgs, cnt, noise = heldout()
    orc = O.AIROracle(params=params, annealing_schedules=O.DEFAULT_ANNEALING, train=False)
    orc.global_step = step
    with torch.no_grad():
        out = orc.forward(imgs, cnt, noise)
    vis = ab.visualize_reconstructions(imgs[:n], out["reconstruction"][:n], out["rec_st_back"][:n],
                                       out["rec_num_digits"][:n], transformer=O.transformer).numpy()
    rows = [np.concatenate(list(vis[r * cols:(r + 1) * cols]), axis=1) for r in range(n // cols)]
    Image.fromarray((np.concatenate(rows, axis=0) * 255).astype(np.uint8)).save(out_png)
    acc = (out["rec_num_digits"] == cnt).float().mean().item()
    print(f"{out_png}: {n} held-out canvases, digit-count accuracy on all {len(cnt)}: {acc:.4f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

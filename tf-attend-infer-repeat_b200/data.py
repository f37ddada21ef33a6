"""Synthetic multi-digit canvases for benchmarks (stand-in for multi_mnist.py:82-183, which
needs the MNIST download): 0..max_digits stroke-like blobs per canvas, uniform placement with
pixel-overlap rejection, background exactly 0.0, digit count uniform over {0..max_digits}."""
from __future__ import annotations

import numpy as np
import torch


def synthetic_canvases(B, canvas_size=50, max_digits=2, seed=0, n_templates=256):
    """-> (images [B, canvas_size**2] float32 in [0,1], counts [B] int32), on the CPU."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
    templates = []
    for _ in range(n_templates):
        hh, ww = rng.randint(14, 25), rng.randint(10, 25)
        cy, cx = (hh - 1) / 2.0, (ww - 1) / 2.0
        r = np.sqrt(((yy[:hh, :ww] - cy) / max(hh / 2.0 - 1.5, 2.0)) ** 2 + ((xx[:hh, :ww] - cx) / max(ww / 2.0 - 1.5, 2.0)) ** 2)
        ring = np.clip(1.0 - np.abs(r - 0.8) * 3.0, 0.0, 1.0)
        bar = np.clip(1.0 - np.abs(xx[:hh, :ww] - cx - rng.uniform(-2, 2)) / 1.6, 0.0, 1.0) * rng.randint(0, 2)
        blob = np.maximum(ring, bar).astype(np.float32)
        blob[blob < 0.15] = 0.0
        templates.append(blob)
    imgs = np.zeros((B, canvas_size, canvas_size), np.float32)
    counts = rng.randint(0, max_digits + 1, size=B).astype(np.int32)
    for b in range(B):
        for _ in range(int(counts[b])):
            for _attempt in range(20):
                blob = templates[rng.randint(n_templates)]
                hh, ww = blob.shape
                top, left = rng.randint(0, canvas_size - hh + 1), rng.randint(0, canvas_size - ww + 1)
                region = imgs[b, top:top + hh, left:left + ww]
                if np.any((region > 0) & (blob > 0)):
                    continue
                np.maximum(region, blob, out=region)
                break
    return torch.from_numpy(imgs.reshape(B, -1)), torch.from_numpy(counts)


def device_canvases(B, seed=0, first_index=0, canvas_size=50, max_digits=2, device="cuda", out=None, with_boxes=False):
    """Multi-digit canvases generated ON the GPU (air_synth_canvases): -> (images [B, cs*cs] fp32, counts [B] int32).
    Image ``first_index + b`` depends only on (seed, first_index + b): data-parallel ranks call this with their own
    ``first_index`` and together hold one global, reproducible data set.  ``out=(images, counts)`` refills buffers
    in place (e.g. AIRModel.input_images / target_num_digits) with no host round trip."""
    from . import ops
    if out is None:
        out = (torch.empty(B, canvas_size * canvas_size, device=device), torch.empty(B, device=device, dtype=torch.int32))
    if with_boxes:   # + the generator's bookkeeping (multi_mnist.py:165-166): positions (x, y), boxes (w, h)
        pos = torch.empty(B, max_digits, 2, device=out[0].device, dtype=torch.int32)
        box = torch.empty_like(pos)
        ops.synth_canvases(out[0], out[1], seed, first_index, canvas_size, max_digits, positions=pos, boxes=box)
        return out[0], out[1], pos, box
    return ops.synth_canvases(out[0], out[1], seed, first_index, canvas_size, max_digits)


# training.py:100-122 -- the configuration the README / checkpoint call "default"
TRAINING_HYPER = dict(
    max_steps=3, max_digits=2, rnn_units=256, canvas_size=50, windows_size=28,
    vae_latent_dimensions=50, vae_recognition_units=(512, 256), vae_generative_units=(256, 512),
    scale_prior_mean=-1.0, scale_prior_variance=0.05, shift_prior_mean=0.0, shift_prior_variance=1.0,
    vae_prior_mean=0.0, vae_prior_variance=1.0, vae_likelihood_std=0.3,
    scale_hidden_units=64, shift_hidden_units=64, z_pres_hidden_units=64,
    z_pres_prior_log_odds=-0.01, z_pres_temperature=1.0, stopping_threshold=0.99,
    learning_rate=1e-4, gradient_clipping_norm=1.0, cnn=False,
)
TRAINING_ANNEALING = {"z_pres_prior_log_odds": {"init": 10000.0, "min": 0.000000001, "factor": 0.1, "iters": 3000,
                                                "staircase": False, "log": True}}

# dense MACs per image: canonical = every step recomputes x @ K; executed = hoisted
MAC_FWD_STEP = 2822144 + 81920 + 448 + 1103360
MAC_BWD_STEP = 5455744
def train_flops_per_image(T=3):
    return 2 * T * (MAC_FWD_STEP + MAC_BWD_STEP)

"""Inference adapter with the contract of the reference's demo/model_wrapper.py:14-52.

``ModelWrapper(model).infer(images)`` takes a list of [canvas, canvas] images and returns six lists
(digits, positions, reconstructions, windows, latents, loss), slicing each image's per-step outputs to its
inferred digit count exactly like the reference does after its session.run.  The reference's ``session`` and
``data_placeholder`` arguments are accepted and ignored (the CUDA model is eager).

Also: the per-digit-count / per-step scalar summaries of air_model.py:160-209, 613-625 as plain dictionaries.
"""
from __future__ import annotations

import numpy as np
import torch


class ModelWrapper:

    def __init__(self, model, session=None, data_placeholder=None, canvas_size=50, window_size=28):
        self.model = model
        self.session = session
        self.data_placeholder = data_placeholder
        self.canvas_size = canvas_size
        self.window_size = window_size

    def infer(self, images, noise=None):
        all_digits, all_positions = [], []
        all_windows, all_latents = [], []
        all_reconstructions, all_loss = [], []
        m = self.model
        B = m.batch_size
        flat = np.stack([np.ravel(np.asarray(img, dtype=np.float32)) for img in images]) if len(images) else \
            np.zeros((0, self.canvas_size ** 2), np.float32)
        for start in range(0, len(flat), B):
            chunk = flat[start:start + B]
            n = len(chunk)
            batch = np.zeros((B, flat.shape[1]), np.float32)     # the model has a fixed batch: pad, then slice
            batch[:n] = chunk
            m.feed(torch.from_numpy(batch).to(m.device))
            m.run(noise)
            rec_digits = m.rec_num_digits[:n].cpu().numpy()
            rec_scales = m.rec_scales[:n].cpu().numpy()
            rec_shifts = m.rec_shifts[:n].cpu().numpy()
            reconstructions = m.reconstruction[:n].cpu().numpy()
            rec_windows = m.rec_windows[:n].cpu().numpy()
            rec_latents = m.rec_latents[:n].cpu().numpy()
            rec_loss = m.reconstruction_loss[:n].cpu().numpy()
            for i in range(n):
                digits = int(rec_digits[i])
                positions, windows, latents = [], [], []
                for j in range(digits):
                    positions.append(np.array([rec_scales[i][j][0]] + list(rec_shifts[i][j])))
                    windows.append(np.reshape(rec_windows[i][j], (self.window_size, self.window_size)))
                    latents.append(rec_latents[i][j])
                all_digits.append(digits)
                all_positions.append(np.array(positions))
                all_reconstructions.append(np.reshape(reconstructions[i], (self.canvas_size, self.canvas_size)))
                all_windows.append(np.array(windows))
                all_latents.append(np.array(latents))
                all_loss.append(rec_loss[i])
        return all_digits, all_positions, all_reconstructions, all_windows, all_latents, all_loss


def summarize_by_digit_count(values, digits, max_digits, name):
    """air_model.py:160-182: mean of ``values`` over the images with exactly i digits, and over all."""
    values = torch.as_tensor(values).float().reshape(-1)
    digits = torch.as_tensor(digits).reshape(-1)
    out = {}
    for i in range(max_digits + 1):
        sel = values[digits == i]
        out[f"{name}_{i}_dig"] = sel.mean().item() if sel.numel() else float("nan")
    out[f"{name}_all_dig"] = values.mean().item() if values.numel() else float("nan")
    return out


def summarize_by_step(tensor, steps, target_digits, max_steps, max_digits, name, one_more_step=False, all_steps=False):
    """air_model.py:184-209: per-step summaries restricted to the items that actually executed that step."""
    tensor = torch.as_tensor(tensor).float()
    steps = torch.as_tensor(steps).reshape(-1)
    target_digits = torch.as_tensor(target_digits).reshape(-1)
    if tensor.shape[1] < max_steps:
        tensor = torch.nn.functional.pad(tensor, (0, max_steps - tensor.shape[1]))
    out = {}
    for i in range(max_steps):
        if all_steps:
            out.update(summarize_by_digit_count(tensor[:, i], target_digits, max_digits, f"{name}_{i + 1}_step"))
        else:
            mask = steps > (i - (1 if one_more_step else 0))
            out.update(summarize_by_digit_count(tensor[:, i][mask], target_digits[mask], max_digits, f"{name}_{i + 1}_step"))
    return out


def evaluation_summaries(model, target_num_digits=None):
    """The numeric summaries the reference logs for its test model (air_model.py:613-625)."""
    t = model.target_num_digits if target_num_digits is None else target_num_digits
    t = t.cpu()
    d = model.rec_num_digits.cpu()
    md, ms = model.max_digits, model.max_steps
    out = {}
    out.update(summarize_by_digit_count(d, t, md, "steps"))
    out.update(summarize_by_digit_count(model.reconstruction_loss.cpu(), t, md, "rec_loss"))
    out.update(summarize_by_digit_count((t == d).float(), t, md, "digit_acc"))
    out.update(summarize_by_digit_count(model.loss_per_item.cpu(), t, md, "total_loss"))
    out.update(summarize_by_step(model.rec_scales[:, :, 0].cpu(), d, t, ms, md, "scale"))
    out.update(summarize_by_step(model.z_pres_probs.cpu(), d, t, ms, md, "z_pres_prob", all_steps=True))
    out.update(summarize_by_step(model.z_pres_kls.cpu(), d, t, ms, md, "z_pres_kl", one_more_step=True))
    out.update(summarize_by_step(model.scale_kls.cpu(), d, t, ms, md, "scale_kl"))
    out.update(summarize_by_step(model.shift_kls.cpu(), d, t, ms, md, "shift_kl"))
    out.update(summarize_by_step(model.vae_kls.cpu(), d, t, ms, md, "vae_kl"))
    return out

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
        all_digits, all_positions, all_reconstructions, all_windows, all_latents, all_loss = ([] for _ in range(6))
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
            fetched = {k: getattr(m, k)[:n].cpu().numpy() for k in
                       ("rec_num_digits", "rec_scales", "rec_shifts", "reconstruction", "rec_windows", "rec_latents",
                        "reconstruction_loss")}
            ws, cs = self.window_size, self.canvas_size
            pose = np.concatenate([fetched["rec_scales"], fetched["rec_shifts"]], axis=2)      # [n, T, 3] = (s, x, y)
            for i, d in enumerate(int(v) for v in fetched["rec_num_digits"]):
                # per image: only the first ``d`` (= inferred digit count) steps are reported, as the reference does
                empty = np.array([])
                all_digits.append(d)
                all_positions.append(pose[i, :d].copy() if d else empty)
                all_windows.append(fetched["rec_windows"][i, :d].reshape(d, ws, ws) if d else empty)
                all_latents.append(fetched["rec_latents"][i, :d].copy() if d else empty)
                all_reconstructions.append(fetched["reconstruction"][i].reshape(cs, cs))
                all_loss.append(fetched["reconstruction_loss"][i])
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


def evaluation_summaries(model, target_num_digits=None, reference_trip_count=True):
    """The numeric summaries the reference logs for its test model (air_model.py:613-625).

    The reference's while_loop leaves as soon as the WHOLE batch has stopped (air_model.py:271-275) and
    _summarize_by_step zero-pads the steps that never ran (air_model.py:188); this model always runs max_steps
    iterations, so with ``reference_trip_count`` the columns at and beyond ``model.executed_steps`` are zeroed first
    (one host sync) -- that is what the reference would have logged for a batch that stops early."""
    t = model.target_num_digits if target_num_digits is None else target_num_digits
    t = t.cpu()
    d = model.rec_num_digits.cpu()
    md, ms = model.max_digits, model.max_steps
    # (stand-ins that already carry the reference's dynamic time dimension have no executed_steps: nothing to trim)
    n_exec = getattr(model, "executed_steps", ms) if reference_trip_count else ms

    def per_step(x):
        x = x.cpu().clone()
        x[:, n_exec:] = 0
        return x
    out = {}
    out.update(summarize_by_digit_count(d, t, md, "steps"))
    out.update(summarize_by_digit_count(model.reconstruction_loss.cpu(), t, md, "rec_loss"))
    out.update(summarize_by_digit_count((t == d).float(), t, md, "digit_acc"))
    out.update(summarize_by_digit_count(model.loss_per_item.cpu(), t, md, "total_loss"))
    out.update(summarize_by_step(per_step(model.rec_scales[:, :, 0]), d, t, ms, md, "scale"))
    out.update(summarize_by_step(per_step(model.z_pres_probs), d, t, ms, md, "z_pres_prob", all_steps=True))
    out.update(summarize_by_step(per_step(model.z_pres_kls), d, t, ms, md, "z_pres_kl", one_more_step=True))
    out.update(summarize_by_step(per_step(model.scale_kls), d, t, ms, md, "scale_kl"))
    out.update(summarize_by_step(per_step(model.shift_kls), d, t, ms, md, "shift_kl"))
    out.update(summarize_by_step(per_step(model.vae_kls), d, t, ms, md, "vae_kl"))
    return out

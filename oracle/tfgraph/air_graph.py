"""Drive the reference's serialized training / test graph (model/air-model.meta) through the numpy interpreter.
TEST INFRASTRUCTURE.  Variable, noise and feed naming of that graph:

* variables ``air/rnn/<name>`` (36 trainables), Adam slots ``air/training/air/rnn/<name>/Adam{,_1}``,
  ``air/training/beta{1,2}_power``, ``air/global_step``;
* the train model reads ``pipeline/shuffle_batch:0`` (images [B,2500]) and ``:1`` (digit counts, int32 [B]);
  the test model ``air_1`` (train=False, reuse=True: same variables) reads ``pipeline/Placeholder{,_1}``;
* per-iteration noise ops inside ``<scope>/rnn/while``: ``scale``, ``shift``, ``vae/rec_sample``, ``vae/gen_sample``
  (RandomStandardNormal) and ``z_pres/gumbel`` (RandomUniform).
"""
from __future__ import annotations

import numpy as np

from .interp import Interpreter, run_in_big_stack

META = "/root/reference/model/air-model.meta"
_NOISE = {"scale/random_normal/RandomStandardNormal": "scale", "shift/random_normal/RandomStandardNormal": "shift",
          "vae/rec_sample/random_normal/RandomStandardNormal": "vae_latent",
          "vae/gen_sample/random_normal/RandomStandardNormal": "vae_like",
          "z_pres/gumbel/random_uniform/RandomUniform": "concrete_u"}


def _np(v):
    return v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)


def graph_variables(params, adam_m=None, adam_v=None, global_step=0, beta1_power=0.9, beta2_power=0.999):
    out = {"air/global_step": np.int32(global_step), "air/training/beta1_power": np.float32(beta1_power),
           "air/training/beta2_power": np.float32(beta2_power)}
    for k, v in params.items():
        v = _np(v).astype(np.float32)
        out["air/rnn/" + k] = v
        out[f"air/training/air/rnn/{k}/Adam"] = _np(adam_m[k]) if adam_m else np.zeros_like(v)
        out[f"air/training/air/rnn/{k}/Adam_1"] = _np(adam_v[k]) if adam_v else np.zeros_like(v)
    return out


def _random_fn(noise, used):
    def fn(node, it, shape):
        for suffix, key in _NOISE.items():
            if node.endswith("/rnn/while/" + suffix):
                v = _np(noise[key])[it]
                used.add((key, it))
                return v.reshape(shape)
        raise KeyError(f"no injected noise for {node}")
    return fn


PER_STEP = {"rec_scales": "rec_scales", "rec_shifts": "rec_shifts", "rec_st_back": "rec_st_back",
            "rec_windows": "rec_windows", "rec_latents": "rec_windows_1", "z_pres_probs": "z_pres_probs",
            "z_pres_kls": "z_pres_kls", "scale_kls": "scale_kls", "shift_kls": "shift_kls", "vae_kls": "vae_kls"}


def _common_fetches(I, scope, nodes):
    out = {"rec_num_digits": I.fetch(f"{scope}/rnn/while/Exit_6"),        # loop variables of air_model.py:548-553
           "canvas_raw": I.fetch(f"{scope}/rnn/while/Exit_4"),
           "running_loss": I.fetch(f"{scope}/rnn/while/Exit_5"),
           "stopping_sum": I.fetch(f"{scope}/rnn/while/Exit_1"),
           "reconstruction": I.fetch(f"{scope}/loss/reconstruction/clipped_rec"),
           "reconstruction_loss": I.fetch(f"{scope}/loss/reconstruction/Neg"),
           "loss_per_item": I.fetch(f"{scope}/add_1")}
    # per-step stop mask ``stopping_sum_new < threshold`` (air_model.py:427-439): an in-loop tensor, one per iteration
    trip = I.trip_count(f"{scope}/rnn/while/{scope}/rnn/while/")
    out["stop_masks"] = np.stack([I.eval(f"{scope}/rnn/while/canvas/Less", 0, t) for t in range(trip)], 1)
    for key, node in PER_STEP.items():
        if f"{scope}/{node}" in nodes:
            out[key] = I.fetch(f"{scope}/{node}")
    return out


def scalar_summaries(I, scope):
    """Every tf.summary.scalar of the model (air_model.py:160-209, 613-625): tag (scope prefix stripped) -> value."""
    out = {}
    with np.errstate(all="ignore"):                        # means over empty selections are NaN, as in TF
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for nd in I.nodes.values():
                if nd.op == "ScalarSummary" and nd.name.startswith(scope + "/"):
                    tag = I.nodes[nd.inputs[0][0]].attr["value"].ravel()[0].decode()
                    out[tag[len(scope) + len("/summaries/"):]] = float(I.eval(*nd.inputs[1], None))
    return out


GRAD = "air/training/gradients/air/rnn/while/"
# In-loop tensors around the two Spatial Transformers: forward operands (frame "f", iteration = step t) and the
# gradients TF's autodiff graph carries in and out of them (frame "b", iteration j = step T-1-j).
ST_TAPS = {
    "theta": ("air/rnn/while/st_forward/stack_2", "f"), "theta_inv": ("air/rnn/while/st_backward/stack_2", "f"),
    "window_recon": ("air/rnn/while/vae/gen_sample/Sigmoid", "f"),
    "d_crop": (GRAD + "vae/Reshape_grad/Reshape", "b"),                     # d loss / d ST(canvas, theta)
    "d_theta": (GRAD + "st_forward/SpatialTransformer/_transform/Reshape_grad/Reshape", "b"),
    "d_writeback": (GRAD + "canvas/Reshape_grad/Reshape", "b"),              # d loss / d ST(window, theta_inv)
    "d_theta_inv": (GRAD + "st_backward/SpatialTransformer/_transform/Reshape_grad/Reshape", "b"),
    "d_window_recon": (GRAD + "st_backward/Reshape_grad/Reshape", "b"),
}


def run_train_step(nodes, params, images, num_digits, noise, float_dtype=np.float32, taps=None, extra_feeds=None,
                   **var_kwargs):
    """One ``sess.run([model.training, loss, accuracy, ...])`` of the reference's train graph.  Returns a dict with
    loss, accuracy, per-step outputs, raw gradients (inputs of the L2Loss ops of clip_by_global_norm), clipped
    gradients (ApplyAdam inputs), the global norm and the updated variables."""
    def body():
        used = set()
        feeds = {"pipeline/shuffle_batch:0": _np(images).astype(np.float32),
                 "pipeline/shuffle_batch:1": _np(num_digits).astype(np.int32)}
        feeds.update(extra_feeds or {})                    # e.g. a chosen upstream gradient at an ST_TAPS tensor
        I = Interpreter(nodes, graph_variables(params, **var_kwargs), feeds, _random_fn(noise, used), float_dtype)
        N = I.nodes
        out = {"loss": I.fetch("air/summaries/loss"), "accuracy": I.fetch("air/summaries/accuracy")}
        out.update(_common_fetches(I, "air", N))
        raw, clipped = {}, {}
        for nd in N.values():
            if nd.op == "ApplyAdam":
                var = nd.inputs[0][0]
                I.fetch(nd.name)
                clipped[var[len("air/rnn/"):]] = I.eval(*nd.inputs[9], None)
        l2 = sorted((nd for nd in N.values() if nd.op == "L2Loss" and "/global_norm/L2Loss" in nd.name),
                    key=lambda nd: (len(nd.name), nd.name))
        adam_vars = [nd.inputs[0][0] for nd in N.values() if nd.op == "ApplyAdam"]
        assert len(l2) == len(adam_vars) == 36
        for nd, var in zip(l2, adam_vars):                 # compute_gradients order == apply order
            g = I.eval(*nd.inputs[0], None)
            assert g.shape == I.variables[var].shape, (nd.name, var)
            raw[var[len("air/rnn/"):]] = g
        I.fetch("air/training/Adam/update")                # 36 ApplyAdam + the two beta-power Assigns
        I.fetch("air/training/Adam")                       # the train op itself: global_step += 1
        out["global_norm"] = I.fetch("air/training/global_norm/global_norm")
        out["executed_steps"] = I.trip_count("air/rnn/while/air/rnn/while/")
        out["raw_grads"], out["clipped_grads"] = raw, clipped
        T = out["executed_steps"]
        for key, (node, fr) in (taps or {}).items():       # [T, ...] in forward step order
            vals = [I.eval(node, 0, t if fr == "f" else T - 1 - t) for t in range(T)]
            out["tap:" + key] = np.stack(vals)
        out["new_variables"] = dict(I.assigned)
        out["noise_used"] = used
        return out
    return run_in_big_stack(body)


def run_test_model(nodes, params, images, num_digits, noise, float_dtype=np.float32, max_steps=None):
    """The ``air_1`` graph (train=False: tf.round of z_pres, air_model.py:385-386) on fed placeholders.
    ``max_steps`` overrides the one place the loop uses it, the constant of cond()'s ``step < max_steps``
    (air_model.py:271-275), e.g. 5 for the inference configuration; the summaries (unrolled for 3 steps when the
    graph was built) are then skipped."""
    def body():
        used = set()
        feeds = {"pipeline/Placeholder:0": _np(images).astype(np.float32),
                 "pipeline/Placeholder_1:0": _np(num_digits).astype(np.int32)}
        if max_steps is not None:
            feeds["air_1/rnn/while/Less/y:0"] = np.int32(max_steps)
        I = Interpreter(nodes, graph_variables(params), feeds, _random_fn(noise, used), float_dtype)
        out = _common_fetches(I, "air_1", I.nodes)
        if max_steps is not None:
            out["loss"] = I.fetch("air_1/summaries/loss")
            out["accuracy"] = I.fetch("air_1/summaries/accuracy")
            out["executed_steps"] = I.trip_count("air_1/rnn/while/air_1/rnn/while/")
            return out
        out["loss"] = I.fetch("air_1/summaries/loss")
        out["accuracy"] = I.fetch("air_1/summaries/accuracy")
        out["executed_steps"] = I.trip_count("air_1/rnn/while/air_1/rnn/while/")
        out["summaries"] = scalar_summaries(I, "air_1")
        # input of tf.summary.image("reconstruction", ...) (air_model.py:211-267, 627-640): [60, 100, 204, 3]
        out["reconstruction_image"] = I.fetch("air_1/summaries/concat_2")
        return out
    return run_in_big_stack(body)


def digest(a, n=4096):
    """Evenly strided sample (<= ~n elements) + Euclidean norm of a tensor: what the golden files keep of the big
    gradient / parameter tensors."""
    a = np.asarray(a)
    flat = a.reshape(-1)
    return flat[::max(1, flat.size // n)].copy(), np.float64(np.linalg.norm(flat.astype(np.float64)))
